#!/usr/bin/env python
"""One full-grid cell with every optional output on (lags, matchups, first-seen): for the ncu launch list."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from farkle_ii_b200.device import get_engine
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
eng = get_engine(0)
table = eng.to_device(pack_strategies(generate_strategy_grid()[0]))
for rep in range(2):
    res = eng.play_tournament(42, k, 0, 4300, table, lags=(1, 2), matchup_min_observations=3, want_first_seen=True)
torch.cuda.synchronize()
print("matchup groups", len(res.matchup_count))
