#!/usr/bin/env python
"""Cost of the RNG lag diagnostics on top of a full-grid cell: python scripts/diag_time.py K [SHUFFLES]."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from farkle_ii_b200 import rng_diagnostics as rd
from farkle_ii_b200.device import get_engine
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nsh = int(sys.argv[2]) if len(sys.argv) > 2 else 4300
eng = get_engine(0)
table = pack_strategies(generate_strategy_grid()[0])
n_games = nsh * (len(table) // k)


def timed(label, **kw):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = eng.play_tournament(42, k, 0, nsh, table, **kw)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"k={k} {label}: {best*1e3:.2f} ms ({n_games/best/1e6:.1f} Mgames/s)")
    return res


timed("tallies only")
res = timed("+ strategy lags (1,)", lags=(1,))
res = timed("+ strategy lags (1,2,5,10)", lags=(1, 2, 5, 10))
res = timed("+ strategy lags (1,) + matchups >= 3", lags=(1,), matchup_min_observations=3)
t0 = time.perf_counter()
state = rd.StrategyLagState.from_launch((1,), nsh, res.lag_stats, res.lag_edges)
groups = rd.MatchupLagGroups.from_launch((1,), res)
t1 = time.perf_counter()
print(f"eligible matchup groups: {len(groups)} of {n_games} games; D2H + canonical order {1e3*(t1-t0):.1f} ms")
t0 = time.perf_counter()
(mask,) = rd.select_matchup_groups([groups], 12)
t1 = time.perf_counter()
print(f"blake2b ids + priority cap -> {int(mask.sum())} groups: {1e3*(t1-t0):.1f} ms")
t0 = time.perf_counter()
rows = state.rows(range(len(table)), k) + groups.rows(12, mask)
t1 = time.perf_counter()
est = sum(r["estimability_status"] == "estimated" for r in rows)
print(f"report rows: {len(rows)} ({est} estimated) in {1e3*(t1-t0):.1f} ms")
ac = np.array([r["autocorr"] for r in rows if r["autocorr"] is not None and r["summary_level"] == "strategy"])
print(f"strategy autocorr: mean {ac.mean():+.5f}, max |.| {np.abs(ac).max():.4f}, band +-{1.96/ (nsh-1)**0.5:.4f}")
