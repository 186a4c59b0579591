#!/usr/bin/env python
"""The reference's whole H2H execution stage (`execute_h2h_schedule`, analysis/h2h_schedule.py:1597)
on a power-planned schedule at production block size, timed twice on this box: with the reference's
own `_simulate_block` on all host cores (its `process_map` pool), and with this repo's
`BatchedBlockRunner` (every pending block advanced per launch on the GPU, behind the stage's
unmodified one-block-at-a-time loop).  Both runs write the stage's full artifact set (block Parquets,
sidecars, execution state, order counts); the script checks that they are byte-identical.

    python scripts/h2h_stage_time.py [N_CANDIDATES]        # default 8 -> 28 pairs x 2 roots x 2 orders

Needs the staged reference (scripts/stage_reference.sh) and git.  Prints one JSON line.
"""
from __future__ import annotations

import json
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
from refpath import reference_root  # noqa: E402

DRIVER = r'''
import os, sys, time, json
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
ref_src, repo, out, kind, n_cand = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
sys.path.insert(0, ref_src)
from pathlib import Path
import pandas as pd, pyarrow as pa, pyarrow.parquet as pq
from farkle.analysis.h2h_schedule import execute_h2h_schedule, plan_h2h_schedule
from farkle.config import AppConfig, ArtifactScope, IOConfig, SimConfig
from farkle.simulation.simulation import generate_strategy_grid
from farkle.simulation.strategies import build_strategy_manifest
from farkle.utils.artifact_contract import make_artifact_sidecar
from farkle.utils.artifacts import write_json_artifact_atomic, write_parquet_artifact_atomic
roots = (11, 22)
cfg = AppConfig(io=IOConfig(results_dir_prefix=Path(out) / "results"),
                sim=SimConfig(seed=roots[0], seed_list=list(roots), n_players_list=[2, 4]))
cfg.screening.practical_delta_by_k = {2: 0.03, 4: 0.03}
cfg.screening.delta_across_k = 0.03
cfg.head2head.total_game_cap = None
cfg.head2head.n_jobs = 1       # identical configuration (it is hashed into every sidecar) in both runs
cfg.resources.scheduler_memory_budget_mb = 8192
cfg.resources.process_tree_warning_threshold_mb = 24576
cfg.resources.aggregate_memory_hard_limit_mb = 32768
cfg.resources.minimum_system_available_memory_mb = 256
cfg.resources.logical_cpu_budget = os.cpu_count() or 1
step = 5160 // n_cand
candidates = tuple(17 + i * step for i in range(n_cand))
family_hash = "a" * 64
membership = pd.DataFrame({"strategy": list(candidates), "final_family": [True] * n_cand,
                           "family_hash": [family_hash] * n_cand})
membership["strategy"] = pd.array(membership["strategy"].tolist(), dtype="Int32")
manifest = {"family_hash": family_hash, "candidates": list(candidates), "candidate_count": n_cand,
            "root_seeds": list(roots), "single_root_execution": False}
common = dict(producer="test", scope=ArtifactScope.H2H_2P, source_scope=ArtifactScope.CROSS_SEED,
              operation="candidate_family_freeze", player_counts=[2], required_player_counts=[2],
              missing_cell_policy="fail", seed_scope="both_roots_combined")
mp = cfg.h2h_candidate_family_path()
write_parquet_artifact_atomic(pa.Table.from_pandas(membership, preserve_index=False), mp,
    sidecar=make_artifact_sidecar(cfg, mp, consistency_columns=membership.columns.tolist(), **common))
jp = cfg.h2h_candidate_family_manifest_path()
write_json_artifact_atomic(manifest, jp, sidecar=make_artifact_sidecar(cfg, jp, consistency_columns=list(manifest), **common))
plan_h2h_schedule(cfg)
sm = cfg.strategy_manifest_root_path()
sm.parent.mkdir(parents=True, exist_ok=True)
build_strategy_manifest(generate_strategy_grid()[0]).to_parquet(sm)
schedule = pq.read_table(cfg.h2h_block_manifest_path()).to_pandas()
kw = dict(n_jobs=os.cpu_count() or 1)        # the reference's pool: all host cores
if kind == "batched":
    sys.path.insert(0, repo)
    from farkle_ii_b200 import h2h
    h2h.simulate_blocks(schedule.to_dict(orient="records")[:1], pd.read_parquet(sm), 1)      # context + kernels warm
    kw = dict(n_jobs=1, block_runner=h2h.BatchedBlockRunner(schedule.to_dict(orient="records"), chunk_games=5000))
t0 = time.perf_counter()
art = execute_h2h_schedule(cfg, chunk_games=5000, **kw)
dt = time.perf_counter() - t0
counts = pq.read_table(art.order_counts).to_pandas()
print("RESULT " + json.dumps({"kind": kind, "seconds": dt, "blocks": len(schedule),
      "n_completed_required": int(schedule["n_completed_required"].iloc[0]),
      "max_attempts": int(schedule["max_attempts"].iloc[0]),
      "games_attempted": int(counts["games_attempted"].sum()), "games_completed": int(counts["games_completed"].sum()),
      "workers": (os.cpu_count() or 1) if kind == "reference" else 1}))
'''


def main() -> None:
    n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ref_root = reference_root()
    if ref_root is None or shutil.which("git") is None:
        print(json.dumps({"unavailable": "needs the staged reference (scripts/stage_reference.sh) and git"}))
        return
    tmp = Path(tempfile.mkdtemp(prefix="fb_h2h_stage_"))
    try:
        ref = tmp / "ref"
        ref.mkdir()
        for name in ("src", "configs", "pyproject.toml"):
            src = ref_root / name
            (shutil.copytree if src.is_dir() else shutil.copy)(src, ref / name)
        (ref / "drive.py").write_text(DRIVER)
        git = ["git", "-c", "user.email=t@example.org", "-c", "user.name=t"]
        for args in (["init", "-q"], ["add", "-A"], ["commit", "-qm", "reference copy"]):
            subprocess.run([*git, *args], cwd=ref, check=True)
        res = {}
        for kind in ("batched", "reference"):
            t0 = time.perf_counter()
            done = subprocess.run([sys.executable, str(ref / "drive.py"), str(ref / "src"), str(ROOT), str(tmp / kind),
                                   kind, str(n_cand)], cwd=ref, capture_output=True, text=True, timeout=3000)
            if done.returncode != 0:
                print(json.dumps({"error": kind, "stderr": done.stderr[-1500:]}))
                return
            line = next(ln for ln in done.stdout.splitlines() if ln.startswith("RESULT "))
            res[kind] = json.loads(line[7:])
            res[kind]["process_seconds"] = time.perf_counter() - t0
        a, b = tmp / "batched", tmp / "reference"
        files = sorted(p.relative_to(a) for p in a.rglob("*") if p.is_file())
        same = files == sorted(p.relative_to(b) for p in b.rglob("*") if p.is_file()) and all(
            (a / f).read_bytes() == (b / f).read_bytes() for f in files)
        print(json.dumps({"stage": "execute_h2h_schedule (analysis/h2h_schedule.py:1597), unmodified",
                          "candidates": n_cand, "artifacts_byte_identical": same, "files": len(files),
                          "gpu_batched_runner": res["batched"], "reference_pool": res["reference"],
                          "stage_speedup": res["reference"]["seconds"] / res["batched"]["seconds"]}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
