#!/usr/bin/env python
"""Kernel times of k=2 / k=4 cells played back to back (as bench.py does) and alone."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from farkle_ii_b200.device import get_engine
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies
eng = get_engine(0)
table = eng.to_device(pack_strategies(generate_strategy_grid()[0]))
def run(seq, label):
    out = []
    for i, k in enumerate(seq):
        eng.play_tournament(42 + (i // 2) % 2, k, 0, 4300, table)
    torch.cuda.synchronize()
    hist = eng.play_kernel_ms_history(len(seq))[::-1]
    print(label, " ".join(f"k{k}:{ms:.2f}" for k, ms in zip(seq, hist)))
run([2, 4] * 4, "alternating, no sync between launches:")
run([4] * 6, "k=4 only:")
run([2] * 4, "k=2 only:")
run([4, 2] * 4, "alternating, k=4 first:")
# exactly bench.py's step: preallocated tallies, zero_() between launches, roots alternate per step
from farkle_ii_b200.layout import TALLY_WIDTH, TOTALS_WIDTH
n = table.numel() // 8
tallies = {k: torch.zeros((1, n, TALLY_WIDTH), dtype=torch.int64, device=eng.device) for k in (2, 4)}
totals = {k: torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=eng.device) for k in (2, 4)}
for i in range(6):
    for k in (2, 4):
        tallies[k].zero_(); totals[k].zero_()
        eng.play_tournament((42, 43)[i % 2], k, 0, 4300, table, tallies=tallies[k], totals=totals[k])
torch.cuda.synchronize()
hist = eng.play_kernel_ms_history(12)[::-1]
print("bench-style loop:", " ".join(f"{ms:.2f}" for ms in hist))
import bench
tb = eng.to_device(bench.full_grid_table())
print("same table as generate_strategy_grid:", bool((tb == table).all()))
for i in range(4):
    for k in (2, 4):
        eng.play_tournament((42, 43)[i % 2], k, 0, 4300, tb)
torch.cuda.synchronize()
print("bench table:", " ".join(f"{ms:.2f}" for ms in eng.play_kernel_ms_history(8)[::-1]))
