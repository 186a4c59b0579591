#!/usr/bin/env python
"""Rows-mode timing of one full-grid cell: python scripts/rows_time.py K [SHUFFLES]."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from farkle_ii_b200.device import get_engine
from farkle_ii_b200.layout import row_dtype
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nsh = int(sys.argv[2]) if len(sys.argv) > 2 else 4300
eng = get_engine(0)
table = pack_strategies(generate_strategy_grid()[0])
n_games = nsh * (len(table) // k)
pin = torch.empty(n_games * row_dtype(k).itemsize, dtype=torch.uint8).pin_memory()
rows_host = pin.numpy().view(row_dtype(k))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t, tot, rows = eng.run_tournament_host(42 + rep, k, 0, nsh, table, want_rows=True, out_rows=rows_host)
    dt = time.perf_counter() - t0
    print(f"k={k} rows to pinned host: {dt*1e3:.1f} ms, {n_games/dt/1e6:.1f} Mgames/s, {rows.nbytes/dt/1e9:.1f} GB/s rows, "
          f"play_kernel {eng.last_play_kernel_ms():.2f} ms, safety {tot[2]}")
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t, tot, _ = eng.run_tournament_host(42 + rep, k, 0, nsh, table)
    dt = time.perf_counter() - t0
    print(f"k={k} tallies only: {dt*1e3:.1f} ms, {n_games/dt/1e6:.1f} Mgames/s")
