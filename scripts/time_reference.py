#!/usr/bin/env python
"""Time the UNMODIFIED reference's own tournament path on this box's host cores (bench.py's
`cpu_baseline_reference` leg and `--impl reference` arm; SURVEY.md section 8d recipe).

    python scripts/time_reference.py --k 2 --root 42 --shuffles 8 --n-jobs 0 [--warm]

Calls `farkle.simulation.run_tournament.run_tournament(config=TournamentConfig(n_players=k, ...),
strategies=<full 5,160 grid>, global_seed=root, n_jobs=J, collect_metrics=True,
row_output_directory=None, num_shuffles=S)` -- the library pool path behind `farkle run`
(simulation/run_tournament.py:1050-1073 over utils/parallel.py:842-1143), no artifact-v3 overhead --
from the checkout `tests/refpath.py` locates (the staged copy baseline/_ref/ on the GPU box), in a
process of its own: no CUDA, nothing of farkle_ii_b200 on the path.  One deterministic batch =
one shuffle, so the pool gets one task per shuffle.  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
from refpath import numba_cache_env, reference_root  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--root", type=int, default=42)
    ap.add_argument("--shuffles", type=int, default=8)
    ap.add_argument("--n-jobs", type=int, default=0, help="0 = os.cpu_count()")
    ap.add_argument("--warm", action="store_true", help="one untimed 1-shuffle run first (Numba JIT)")
    args = ap.parse_args()
    ref = reference_root()
    if ref is None:
        print(json.dumps({"unavailable": "no reference checkout (baseline/_ref not staged)"}))
        return
    numba_cache_env()
    sys.path.insert(0, str(ref / "src"))
    from farkle.simulation import run_tournament as rt
    from farkle.simulation.simulation import generate_strategy_grid
    from farkle.utils import parallel

    strategies = generate_strategy_grid()[0]
    jobs = args.n_jobs if args.n_jobs > 0 else (os.cpu_count() or 1)

    def run(n_shuffles: int, n_jobs: int, out: Path) -> dict:
        cfg = rt.TournamentConfig(n_players=args.k, num_shuffles=n_shuffles, deterministic_batch_size=1)
        guard = parallel.ProcessTreeMemoryGuard(1 << 20, rss_warning_mb=1 << 20,
                                                minimum_system_available_memory_mb=64)
        t0 = time.perf_counter()
        rt.run_tournament(config=cfg, global_seed=args.root, checkpoint_path=out / "ckpt.pkl", n_jobs=n_jobs,
                          collect_metrics=True, row_output_directory=None, num_shuffles=n_shuffles,
                          strategies=strategies, resume=False, memory_guard=guard,
                          write_workload_plan_artifact=False)
        dt = time.perf_counter() - t0
        ck = pickle.loads((out / "ckpt.pkl").read_bytes())
        return {"seconds": dt, "games": int(ck["outcome_counts"]["games_attempted"]),
                "wins_total": int(sum(ck["win_totals"].values()))}

    with tempfile.TemporaryDirectory() as td:
        if args.warm:
            run(1, 1, Path(td) / "warm")
        res = run(args.shuffles, jobs, Path(td) / "run")
    print(json.dumps({"impl": "Isaac-McPadden/Farkle_II run_tournament.run_tournament (unmodified, Python + Numba)",
                      "k": args.k, "root": args.root, "shuffles": args.shuffles, "n_jobs": jobs,
                      "host_cores": os.cpu_count(), "games": res["games"], "seconds": res["seconds"],
                      "games_per_s": res["games"] / res["seconds"], "wins_total": res["wins_total"],
                      "reference": str(ref)}))


if __name__ == "__main__":
    main()
