#!/usr/bin/env python
"""configs[4] across the GPUs of one box: every (root 102, k) cell of the mega config is split over
the ranks by deterministic batch (run_tournament.run_cell: one launch per rank and cell, one NCCL
all-reduce of the tallies per cell).  Strong scaling of one root; launched with torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29540 scripts/mega_multi.py > gpurun_out/mega_N.json
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from farkle_ii_b200 import run_tournament as frt  # noqa: E402
from farkle_ii_b200.device import get_engine  # noqa: E402
from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies  # noqa: E402

saved = os.dup(1)
os.dup2(2, 1)                               # NCCL banners go to stderr
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
table = eng.to_device(pack_strategies(generate_strategy_grid()[0]))
n = table.numel() // 8
ks = (2, 3, 4, 5, 6, 8, 10, 12)
launch = frt._engine_launch(eng, {})


def root_pass(root):
    out = {}
    for k in ks:
        tallies, totals, seen = frt.run_cell(root, k, 4300, table, batch_size=43, launch=launch, rank=rank,
                                             world=world, want_first_seen=True)
        out[k] = (tallies, totals)
    return out


root_pass(101)                              # warm-up: workspaces, NCCL communicators
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for r in range(reps):
    res = root_pass(102)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
games = sum(4300 * (n // k) for k in ks)
if rank == 0:
    tot = {k: res[k][1].cpu().tolist() for k in ks}
    line = {"config": "mega root 102, k in (2,3,4,5,6,8,10,12), 4,300 shuffles per k, full grid",
            "n_gpus": world, "games": games, "ms_per_root": float(ms.item()),
            "games_per_s": games / float(ms.item()) * 1e3,
            "games_attempted_check": sum(t[0] for t in tot.values()),
            "wins_k2_first5": res[2][0][0, :5, 0].cpu().tolist()}
    os.write(saved, (json.dumps(line) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
