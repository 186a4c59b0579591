#!/usr/bin/env python
"""Multi-GPU parity of the Python surface: `run_tournament()` under torchrun must write the same
checkpoint bytes and metrics table as a single process.

    python scripts/check_multigpu.py single /tmp/one          # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29560 scripts/check_multigpu.py multi /tmp/many
    python scripts/check_multigpu.py compare /tmp/one /tmp/many
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
mode, out = sys.argv[1], Path(sys.argv[2])
CELLS = ((42, 2, 430), (42, 4, 301), (7, 6, 130))          # (root, k, shuffles): ragged last batches too

if mode == "compare":
    other = Path(sys.argv[3])
    bad = 0
    for root, k, _ in CELLS:
        for name in (f"{root}_{k}p_checkpoint.pkl", f"{root}_{k}p/{k}p_metrics.parquet"):
            a, b = (out / name).read_bytes(), (other / name).read_bytes()
            same = a == b
            bad += not same
            print(f"{name}: {'identical' if same else 'DIFFERENT'} ({len(a)} bytes)")
    sys.exit(1 if bad else 0)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from farkle_ii_b200 import run_tournament as frt  # noqa: E402
from farkle_ii_b200.strategies import generate_strategy_grid  # noqa: E402

os.dup2(2, 1)
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if mode == "multi":
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
strategies = generate_strategy_grid()[0]
for root, k, shuffles in CELLS:
    cfg = frt.TournamentConfig(n_players=k, num_shuffles=shuffles, deterministic_batch_size=43)
    cell_dir = out / f"{root}_{k}p"
    cell_dir.mkdir(parents=True, exist_ok=True)
    frt.run_tournament(config=cfg, global_seed=root, checkpoint_path=cell_dir / f"{k}p_checkpoint.pkl",
                       collect_metrics=True, num_shuffles=shuffles, strategies=strategies, device=local)
    if rank == 0:
        (out / f"{root}_{k}p_checkpoint.pkl").write_bytes((cell_dir / f"{k}p_checkpoint.pkl").read_bytes())
        print(f"rank 0 wrote cell root={root} k={k} ({world} ranks)", file=sys.stderr)
if mode == "multi":
    dist.destroy_process_group()
