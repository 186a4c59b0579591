#!/usr/bin/env python
"""Run one tournament cell (for ncu / timing): python scripts/profile_cell.py K SHUFFLES [REPS]."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from farkle_ii_b200.device import get_engine  # noqa: E402
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
shuffles = int(sys.argv[2]) if len(sys.argv) > 2 else 430
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rows = len(sys.argv) > 4 and sys.argv[4] == "rows"
eng = get_engine(0)
table = eng.to_device(pack_strategies(generate_strategy_grid()[0]))
for rep in range(reps):
    res = eng.play_tournament(42 + rep, k, 0, shuffles, table, want_rows=rows)
    ms = eng.last_play_kernel_ms()
    tot = res.totals.cpu().numpy()
    print(f"k={k} shuffles={shuffles} games={tot[0]} play_kernel {ms:.3f} ms "
          f"{tot[0] / ms / 1e3:.1f} Mgames/s rolls/game {tot[3] / tot[0]:.1f} "
          f"rolls={tot[3]} dice={tot[4]} rng_words={tot[5]}")
torch.cuda.synchronize()
