#!/usr/bin/env python
"""Per-kernel timeline of bench steps (fb_timeline): python scripts/timeline_step.py [--unpipelined] [STEPS].

Prints, for the last STEPS steps after three warm-up steps, every mark as
`lane name t_ms dt_ms` (lane 1 = the preparation stream) so that the overlap of the preparation of
cell i+1 with the finish / tally passes of cell i can be read off, and the windows between two
play kernels.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
from farkle_ii_b200.device import get_engine  # noqa: E402
from farkle_ii_b200.layout import TALLY_WIDTH, TOTALS_WIDTH  # noqa: E402

unpipelined = "--unpipelined" in sys.argv
nums = [a for a in sys.argv[1:] if a.isdigit()]
steps = int(nums[0]) if nums else 2
eng = get_engine(0)
table = eng.to_device(bench.full_grid_table())
N, K = bench.N_STRATEGIES, bench.CELLS_K
tallies = {k: torch.zeros((1, N, TALLY_WIDTH), dtype=torch.int64, device=eng.device) for k in K}
totals = {k: torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=eng.device) for k in K}


def step(i):
    root, nxt = bench.ROOTS[i % 2], bench.ROOTS[(i + 1) % 2]
    if unpipelined:
        for k in K:
            eng.play_tournament(root, k, 0, bench.SHUFFLES, table, tallies=tallies[k], totals=totals[k])
    else:
        eng.play_cells([(root, k, 0, bench.SHUFFLES, tallies[k], totals[k]) for k in K], table,
                       ahead=(nxt, K[0], 0, bench.SHUFFLES))


for i in range(3):
    step(i)
torch.cuda.synchronize()
eng.timeline(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3, 3 + steps):
    step(i)
e1.record()
torch.cuda.synchronize()
marks = eng.timeline_marks()
eng.timeline(False)
print(f"# {'unpipelined' if unpipelined else 'pipelined'}: {steps} steps, {e0.elapsed_time(e1) / steps:.3f} ms per step")
last = {0: None, 1: None}
for lane, name, ms in marks:
    dt = ms - last[lane] if last[lane] is not None else 0.0
    last[lane] = ms
    print(f"{lane} {name:14s} {ms:9.3f}  +{dt:7.3f}")
# windows between consecutive play kernels on the caller's stream
ends = [ms for lane, name, ms in marks if name == "play_kernel"]
begins = [ms for lane, name, ms in marks if name == "play_begin"]
gaps = [b - e for e, b in zip(ends[:-1], begins[1:])]
print("# windows between play kernels (ms):", " ".join(f"{g:.3f}" for g in gaps))
