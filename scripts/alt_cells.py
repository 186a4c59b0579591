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
