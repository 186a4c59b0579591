"""N>1 path on CPU: two `gloo` ranks shard a (root, k) cell by deterministic batch and merge the
tally tensors with the same all-reduce the GPUs use over NCCL (run_tournament.run_cell).

The per-rank launches are served by the oracle here (no GPU in this container); what is under
test is the partition (disjoint coordinate ranges, every shuffle played exactly once), the
ranks-without-work case and the merge.
"""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_launch(root_seed, k, shuffle0, n_shuffles, table, tallies, totals):
    import oracle

    t, tot, _ = oracle.play_tournament(root_seed, k, shuffle0, n_shuffles, table, n_threads=1)
    tt, to = torch.from_numpy(t), torch.from_numpy(tot)
    if tallies is not None:
        tt += tallies
        to += totals
    return tt, to


def _worker(rank: int, world: int, port: int, num_shuffles: int, batch: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from farkle_ii_b200 import run_tournament as frt
        from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

        table = pack_strategies(generate_strategy_grid(
            score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
            consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
            run_up_score_opts=[True])[0])
        launches = []

        def launch(*a):
            launches.append((a[2], a[3]))
            return _oracle_launch(*a)

        tallies, totals = frt.run_cell(42, 2, num_shuffles, table, batch_size=batch, launch=launch,
                                       rank=dist.get_rank(), world=dist.get_world_size())
        # the counters' key order (MIN over ranks) and the lag state (all-gather + join in order)
        from farkle_ii_b200 import rng_diagnostics as rd
        from oracle_engine import OracleEngine

        eng = OracleEngine()

        def launch_seen(root_seed, k, s0, n, tab, tallies, totals):
            res = eng.play_tournament(root_seed, k, s0, n, tab, tallies=tallies, totals=totals,
                                      want_first_seen=True)
            return res.tallies, res.totals, res.first_seen

        _, _, seen = frt.run_cell(42, 2, num_shuffles, table, batch_size=batch, launch=launch_seen,
                                  rank=dist.get_rank(), world=dist.get_world_size(), want_first_seen=True)
        lag = rd.cell_lag_state(42, 2, num_shuffles, table, (1, 3), batch_size=batch, rank=dist.get_rank(),
                                world=dist.get_world_size(), engine=eng) if num_shuffles >= batch else None
        np.savez(Path(out_dir) / f"rank{rank}.npz", tallies=tallies.numpy(), totals=totals.numpy(),
                 launches=np.array(launches, dtype=np.int64).reshape(-1, 2), seen=seen.numpy(),
                 lag_stats=lag.stats if lag else np.zeros(0), lag_tail=lag.tail if lag else np.zeros(0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_shuffles,batch", [(50, 8), (5, 30)])
def test_two_ranks_equal_one(tmp_path, num_shuffles, batch):
    import oracle

    oracle.build()
    mp.spawn(_worker, args=(2, _free_port(), num_shuffles, batch, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in (0, 1))
    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    table = pack_strategies(generate_strategy_grid(
        score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
        consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
        run_up_score_opts=[True])[0])
    want_t, want_tot, _ = oracle.play_tournament(42, 2, 0, num_shuffles, table, n_threads=2)
    for r in (r0, r1):                                   # every rank holds the merged cell
        assert np.array_equal(r["tallies"], want_t)
        assert np.array_equal(r["totals"], want_tot)
    from farkle_ii_b200 import rng_diagnostics as rd
    from oracle_engine import OracleEngine

    whole = OracleEngine().play_tournament(42, 2, 0, num_shuffles, table, want_first_seen=True, lags=())
    want_seen = whole.first_seen.numpy().astype(np.int64)
    want_seen[want_seen < 0] = np.iinfo(np.int64).max
    for r in (r0, r1):
        assert np.array_equal(r["seen"], want_seen)
    if num_shuffles >= batch:
        res = OracleEngine().play_tournament(42, 2, 0, num_shuffles, table, lags=(1, 3))
        want_lag = rd.StrategyLagState.from_launch((1, 3), num_shuffles, res.lag_stats, res.lag_edges)
        for r in (r0, r1):
            assert np.array_equal(r["lag_stats"], want_lag.stats)
            assert np.array_equal(r["lag_tail"], want_lag.tail)
    played = sorted(s for r in (r0, r1) for s0, n in r["launches"] for s in range(s0, s0 + n))
    assert played == list(range(num_shuffles))           # disjoint and complete
    if num_shuffles <= batch:
        assert len(r1["launches"]) == 0                  # a rank without work still reduces


# ------------------------------------------------------------------- several cells, one collective
def _oracle_play_cells(log):
    def play_cells(segments, table):
        import oracle

        for root, k, s0, n, tallies, totals in segments:
            log.append((root, k, s0, n))
            t, tot, _ = oracle.play_tournament(root, k, s0, n, table, n_threads=1)
            tallies += torch.from_numpy(t)
            totals += torch.from_numpy(tot)
    return play_cells


_CELLS = [(5, 2, 40), (5, 4, 40), (6, 5, 23), (6, 2, 7)]


def _cells_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from farkle_ii_b200 import run_tournament as frt
        from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

        table = pack_strategies(generate_strategy_grid(
            score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
            consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
            run_up_score_opts=[True])[0])
        log: list = []
        tallies, totals = frt.run_cells(_CELLS, table, batch_size=6, play_cells=_oracle_play_cells(log),
                                        rank=rank, world=world)
        np.savez(Path(out_dir) / f"cells{rank}.npz", tallies=tallies.numpy(), totals=totals.numpy(),
                 log=np.array(log, dtype=np.int64).reshape(-1, 4))
    finally:
        dist.destroy_process_group()


def test_cells_over_two_ranks_one_collective(tmp_path):
    """`run_cells`: the cells of a run dealt to the ranks along one line (whole cells plus at most two
    partial ones per rank, whole deterministic batches), one all-reduce of the stacked tensors:
    every rank ends up with every cell's merged tallies, every shuffle is played exactly once."""
    import oracle

    oracle.build()
    mp.spawn(_cells_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"cells{r}.npz") for r in (0, 1))
    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    table = pack_strategies(generate_strategy_grid(
        score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
        consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
        run_up_score_opts=[True])[0])
    for i, (root, k, shuffles) in enumerate(_CELLS):
        want_t, want_tot, _ = oracle.play_tournament(root, k, 0, shuffles, table, n_threads=2)
        for r in (r0, r1):
            assert np.array_equal(r["tallies"][i], want_t) and np.array_equal(r["totals"][i], want_tot)
    played = sorted((int(root), int(k), s) for r in (r0, r1) for root, k, s0, n in r["log"]
                    for s in range(s0, s0 + n))
    assert played == sorted((root, k, s) for root, k, shuffles in _CELLS for s in range(shuffles))
    assert len(r0["log"]) and len(r1["log"])


def test_plan_cells_properties():
    """The planner alone: complete and disjoint cover in whole batches, balanced estimated load,
    at most two partial cells per rank, identical plan on every call, mega config on 1..8 ranks."""
    from farkle_ii_b200 import run_tournament as frt

    ks = (2, 3, 4, 5, 6, 8, 10, 12)
    for roots in ((102,), (102, 103)):
        cells = [(r, k, 4300) for r in roots for k in ks]
        single = sum(frt.cell_cost_ms(k, 5160, 4300) + frt.SEGMENT_OVERHEAD_MS for _, k, _ in cells)
        for world in (1, 2, 3, 4, 8):
            plan = frt.plan_cells(cells, 5160, world, batch_size=43)
            assert plan == frt.plan_cells(cells, 5160, world, batch_size=43) and len(plan) == world
            cover: dict = {}
            for segs in plan:
                partial = 0
                for sg in segs:
                    assert sg.shuffle0 % 43 == 0 and (sg.n_shuffles % 43 == 0 or sg.shuffle0 + sg.n_shuffles == 4300)
                    assert (sg.root_seed, sg.k) == cells[sg.cell][:2]
                    cover.setdefault(sg.cell, []).append((sg.shuffle0, sg.n_shuffles))
                    partial += sg.n_shuffles != 4300
                assert partial <= 2
            for i in range(len(cells)):
                runs = sorted(cover[i])
                assert runs[0][0] == 0 and sum(n for _, n in runs) == 4300
                assert all(a[0] + a[1] == b[0] for a, b in zip(runs, runs[1:]))
            loads = [sum(frt.SEGMENT_OVERHEAD_MS + frt.cell_cost_ms(sg.k, 5160, sg.n_shuffles) for sg in segs)
                     for segs in plan]
            assert max(loads) <= single / world * 1.04 + 1.0       # within 4 % + 1 ms of perfect balance
    # odd sizes: ragged last batch, more ranks than batches, empty list
    plan = frt.plan_cells([(1, 2, 10), (1, 4, 3)], 80, 5, batch_size=4)
    assert sorted((sg.cell, sg.shuffle0, sg.n_shuffles) for segs in plan for sg in segs) == \
        [(0, 0, 4), (0, 4, 4), (0, 8, 2), (1, 0, 3)]
    assert frt.plan_cells([], 80, 3, batch_size=4) == [[], [], []]


# ------------------------------------------------- bench.py's strong-scaling leg, two gloo ranks
def _strong_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from datetime import timedelta

    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=timedelta(seconds=120))
    try:
        import json
        import time

        import bench
        from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

        table = pack_strategies(generate_strategy_grid(
            score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
            consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
            run_up_score_opts=[True])[0])

        def timer(fn, reps):
            out = fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                out = fn()
            return (time.perf_counter() - t0) * 1e3 / reps, out

        res = bench.strong_scaling_leg(_oracle_play_cells([]), table, rank, world, 2, timer,
                                       cells=[(9, 2, 12), (9, 4, 12), (9, 5, 7)], n_strategies=len(table), batch=3)
        (Path(out_dir) / f"strong{rank}.json").write_text(json.dumps(res))
    finally:
        dist.destroy_process_group()


def test_bench_strong_scaling_leg_two_ranks(tmp_path):
    """The strong-scaling leg of bench.py as the ranks run it (N-rank pass with one all-reduce, then
    rank 0 alone WITHOUT a collective while the others wait at a barrier): completes, both passes
    give identical tallies.  (A collective entered by rank 0 alone deadlocked the 8-GPU run once.)"""
    import json

    import oracle

    oracle.build()
    mp.spawn(_strong_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (json.loads((tmp_path / f"strong{r}.json").read_text()) for r in (0, 1))
    games = 12 * 40 + 12 * 20 + 7 * 16
    for r in (r0, r1):
        assert r["n_gpus"] == 2 and r["games_attempted"] == games and r["identical_tallies_n1_vs_nN"]
        assert sorted(x for segs in r["plan"] for x in segs) == sorted(x for segs in r0["plan"] for x in segs)
    assert r0["n1_ms"] > 0 and r0["nN_ms"] > 0 and len(r0["plan"]) == 2
