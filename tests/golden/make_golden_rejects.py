#!/usr/bin/env python
"""Find tournament games whose dice meet a REJECTED Lemire half (leftover < 4: four 32-bit values out
of 2^32, about seven games of an 11 M-game cell) -> tests/golden/rejects.json.

The CUDA kernel turns halves into queued face codes ahead of the rolls and keeps a rejected half in
the queue as a skip code (csrc/play.cuh, face queue): these games are the ones that exercise that
path.  The list is found with the C oracle (`oracle.scan_rejected_halves`); the oracle itself is
pinned to the unmodified reference by tests/test_oracle_golden.py, and the test that uses this
fixture compares the kernel with the oracle on exactly these games.

    python tests/golden/make_golden_rejects.py        # ~2 min on 8 cores
"""
import json
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies  # noqa: E402

THREADS = 8
CELLS = [(42, 2, 4300), (42, 4, 4300), (43, 3, 2600)]   # (root, k, shuffles scanned)


def main() -> None:
    table = pack_strategies(generate_strategy_grid()[0])
    found = []
    for root, k, n_sh in CELLS:
        step = -(-n_sh // (THREADS * 4))
        spans = [(s0, min(step, n_sh - s0)) for s0 in range(0, n_sh, step)]
        with ThreadPoolExecutor(THREADS) as ex:
            parts = list(ex.map(lambda sp: oracle.scan_rejected_halves(root, k, sp[0], sp[1], table), spans))
        hits = np.concatenate(parts) if parts else np.zeros((0, 3), np.uint64)
        print(f"root {root} k={k}: {len(hits)} games with a rejected half in {n_sh} shuffles")
        for sh, g, r in hits.tolist():
            found.append({"root": root, "k": k, "shuffle": int(sh), "game": int(g), "rejects": int(r)})
    out = ROOT / "tests" / "golden" / "rejects.json"
    out.write_text(json.dumps({"grid": "default full grid (5,160 strategies)", "games": found}, indent=1) + "\n")
    print(f"{len(found)} games -> {out}")


if __name__ == "__main__":
    main()
