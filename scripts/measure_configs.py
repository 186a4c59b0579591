#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (parity-tested elsewhere; bench.py times configs[1]).

    python scripts/measure_configs.py > gpurun_out/configs.json

* configs[2]  6-player games over the full grid (4,300 shuffles, root 42)
* configs[3]  h2h_2p block execution: all pairs of the first 150 strategy ids x 2 orders, root 42,
              n_completed_required 2,191, max_attempts 4,382
* configs[4]  mega config: full grid, k in {2,3,4,5,6,8,10,12}, root 102, 4,300 shuffles per k
Times are CUDA-event times on the launching stream, tallies resident in HBM.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from farkle_ii_b200.device import get_engine  # noqa: E402
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies  # noqa: E402

eng = get_engine(0)
table_host = pack_strategies(generate_strategy_grid()[0])
table = eng.to_device(table_host)
N = len(table_host)
out = {}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


def cell(root, k, shuffles=4300):
    ms, res = timed(lambda: eng.play_tournament(root, k, 0, shuffles, table))
    tot = res.totals.cpu().numpy() // 1  # accumulated once per call (fresh tensors)
    games = shuffles * (N // k)
    return {"k": k, "games": games, "ms": ms, "games_per_s": games / ms * 1e3,
            "play_kernel_ms": eng.last_play_kernel_ms(), "rolls_per_game": float(tot[3] / tot[0]),
            "safety_limit_games": int(tot[2])}


# configs[0]: fast grid (80 strategies), k=2, seed 42, 600 shuffles = 24,000 games (launch-latency bound)
fast_host = pack_strategies(generate_strategy_grid(
    score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
    consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
    run_up_score_opts=[True])[0])
fast = eng.to_device(fast_host)
ms, res = timed(lambda: eng.play_tournament(42, 2, 0, 600, fast), reps=10)
out["fast_config_k2_seed42"] = {"games": 24000, "ms": ms, "games_per_s": 24000 / ms * 1e3,
                                "wins_42_46_51": res.tallies.cpu().numpy()[0][[42, 46, 51], 0].tolist()}

# the Python surface end to end: run_tournament() for the full k=2 cell (checkpoint + metrics parquet)
import tempfile  # noqa: E402

from farkle_ii_b200 import run_tournament as frt  # noqa: E402

with tempfile.TemporaryDirectory() as td:
    cfg = frt.TournamentConfig(n_players=2, num_shuffles=4300, deterministic_batch_size=43)
    kw = dict(config=cfg, global_seed=42, checkpoint_path=Path(td) / "2p_checkpoint.pkl",
              collect_metrics=True, num_shuffles=4300, strategies=generate_strategy_grid()[0])
    frt.run_tournament(**kw)
    t0 = time.perf_counter()
    frt.run_tournament(**kw)
    dt = time.perf_counter() - t0
out["run_tournament_python_surface_k2"] = {
    "games": 4300 * (N // 2), "seconds": dt, "games_per_s": 4300 * (N // 2) / dt,
    "note": "farkle_ii_b200.run_tournament.run_tournament(): strategy packing, launch, tallies D2H, "
            "OutcomeCounter / metric dict building, checkpoint pickle and 2p_metrics.parquet"}

out["k6_full_grid"] = cell(42, 6)
mega = [cell(102, k) for k in (2, 3, 4, 5, 6, 8, 10, 12)]
out["mega_root_102"] = {"cells": mega, "games": sum(c["games"] for c in mega),
                        "ms": sum(c["ms"] for c in mega)}
out["mega_root_102"]["games_per_s"] = out["mega_root_102"]["games"] / out["mega_root_102"]["ms"] * 1e3

# ---- H2H: 150 strategies -> 11,175 pairs x 2 orders = 22,350 blocks
ids = np.arange(150)
a, b = np.triu_indices(150, 1)
pair_id = np.arange(len(a), dtype=np.uint64)
blocks_pair = np.repeat(pair_id, 2)
order = np.tile(np.array([0, 1], dtype=np.uint8), len(a))
s1 = np.where(order == 0, np.repeat(a, 2), np.repeat(b, 2))
s2 = np.where(order == 0, np.repeat(b, 2), np.repeat(a, 2))
seat1, seat2 = table_host[s1], table_host[s2]
target, max_attempts = 2191, 4382
nb = len(blocks_pair)


def h2h():
    prog = np.zeros((nb, 5), dtype=np.int32)
    rounds = attempts = 0
    while True:
        need = np.minimum(np.maximum(target - prog[:, 1], 0), np.maximum(max_attempts - prog[:, 0], 0))
        act = np.flatnonzero(need > 0)
        if len(act) == 0:
            break
        oc, d_na, _, _ = eng.play_h2h(42, blocks_pair[act], order[act], seat1[act], seat2[act],
                                      prog[act, 0].astype(np.uint32), need[act].astype(np.uint32))
        prog[act] = eng.h2h_resolve(d_na, oc, np.full(len(act), target, dtype=np.int32), prog[act])
        rounds += 1
        attempts += int(need[act].sum())
    return prog, rounds, attempts


h2h()
dt = 1e9
for _ in range(3):  # best of three (the first call also sizes the 8 GB workspace)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prog, rounds, attempts = h2h()
    torch.cuda.synchronize()
    dt = min(dt, time.perf_counter() - t0)
out["h2h_2p"] = {"blocks": nb, "n_completed_required": target, "max_attempts": max_attempts,
                 "attempts_played": attempts, "launch_rounds": rounds, "seconds": dt,
                 "attempts_per_s": attempts / dt, "complete_blocks": int((prog[:, 1] >= target).sum()),
                 "unresolved_blocks": int(((prog[:, 1] < target) & (prog[:, 0] >= max_attempts)).sum()),
                 "games_completed": int(prog[:, 1].sum()), "games_safety_limit": int(prog[:, 2].sum()),
                 "note": "wall clock including host progress bookkeeping, H2D block tables and D2H progress"}
# the same schedule through the Python surface (block dicts in, progress dicts out)
from farkle_ii_b200 import h2h as fh2h  # noqa: E402

manifest = fh2h.build_strategy_manifest(generate_strategy_grid()[0])
block_dicts = [{"block_id": f"b{i}", "family_hash": "f", "schedule_hash": "s", "root_seed": 42,
                "pair_id": int(blocks_pair[i]), "order": int(order[i]), "seat1_strategy": int(s1[i]),
                "seat2_strategy": int(s2[i]), "n_completed_required": target, "max_attempts": max_attempts,
                "rng_scheme_version": 2, "rng_purpose_namespace": 203} for i in range(nb)]
fh2h.simulate_blocks(block_dicts[:64], manifest, max_attempts)
t0 = time.perf_counter()
results = fh2h.simulate_blocks(block_dicts, manifest, max_attempts)
dt = time.perf_counter() - t0
out["h2h_2p_python_surface"] = {
    "blocks": nb, "seconds": dt, "attempts": int(sum(r["games_attempted"] for r in results)),
    "attempts_per_s": sum(r["games_attempted"] for r in results) / dt,
    "complete_blocks": sum(r["completion_status"] == "complete" for r in results),
    "note": "farkle_ii_b200.h2h.simulate_blocks: manifest lookup, launches, early-stop resolve, one result "
            "dict per block with the reference's fields (incl. the attempt-range SHA-256)"}
print(json.dumps(out, indent=1))
