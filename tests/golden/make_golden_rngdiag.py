#!/usr/bin/env python
"""Golden output of the reference's WHOLE RNG-diagnostics stage on a tiny run.

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden_rngdiag.py

Copies the unmodified checkout to a scratch directory, git-initialises it (the artifact contract
binds outputs to a code identity), then runs -- through the reference's own entry points --
`simulation.runner.run_tournament(cfg)` (fast grid: 80 strategies, seed 42, k = 2 and 4, the
planner's 12 shuffles, rows on) followed by the analysis stages `ingest`, `curate`, `combine` and
`rng_diagnostics.run(cfg, lags=(1, 2))`, and stores the rows of the resulting
`rng_diagnostics.parquet` (strategy and matchup groups) plus the counts of its summary as
`rng_diagnostics_fast42.json`.  tests/test_rng_diagnostics.py rebuilds the same table from the lag
sums of the tournament reduction.  Only values the reference computes are stored.
"""
from __future__ import annotations

import json
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")

OVERLAY = """
io:
  results_dir_prefix: "{prefix}"
sim:
  n_players_list: [2, 4]
  seed: 42
  seed_list: [42]
  n_jobs: 1
batching:
  target_batches: 2
  min_shuffles_per_batch: 5
screening:
  resolution_delta: 0.5
resources:
  logical_cpu_budget: 2
  scheduler_memory_budget_mb: 2048
  process_tree_warning_threshold_mb: 4096
  aggregate_memory_hard_limit_mb: 6144
  minimum_system_available_memory_mb: 256
  os_memory_limit_enabled: false
  os_memory_limit_required: false
  allow_unenforced_memory_fallback: true
"""

DRIVER = """
import json, os, sys
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.path.insert(0, sys.argv[1])
from pathlib import Path
import pyarrow.parquet as pq
from farkle.config import load_app_config
from farkle.simulation import runner
from farkle.analysis import combine, curate, ingest, rng_diagnostics
cfg = load_app_config(Path(sys.argv[1]).parent / "configs" / "fast_config.yaml", Path(sys.argv[2]))
games = runner.run_tournament(cfg)
for stage in (ingest, curate, combine):
    stage.run(cfg)
rng_diagnostics.run(cfg, lags=(1, 2))
table = pq.read_table(cfg.rng_output_path("rng_diagnostics.parquet"))
summary = json.loads(cfg.rng_output_path("rng_diagnostics_summary.json").read_text())
keep = ("normalized_lags", "minimum_usable_observations", "effective_matchup_group_cap",
        "candidate_strategy_group_count", "candidate_matchup_group_count", "eligible_strategy_group_count",
        "eligible_matchup_group_count", "selected_group_count", "below_minimum_group_count",
        "deterministically_capped_group_count")
from farkle.analysis.combine import combined_partition_paths
with pq.ParquetFile(combined_partition_paths(cfg)[0]) as f:
    seat_columns = [n for n in f.schema_arrow.names if n.startswith("P") and n.endswith("_strategy")]
out = {"games": games, "lags": [1, 2], "root_seed": 42, "ks": [2, 4], "n_strategies": 80,
       "seat_strategy_columns": len(seat_columns),
       "summary": {k: summary[k] for k in keep if k in summary}, "rows": table.to_pylist()}
Path(sys.argv[3]).write_text(json.dumps(out, indent=0, sort_keys=True) + "\\n")
"""


def main() -> None:
    with tempfile.TemporaryDirectory() as td:
        ref = Path(td) / "ref"
        ref.mkdir()
        for name in ("src", "configs", "pyproject.toml"):
            src = REF / name
            (shutil.copytree if src.is_dir() else shutil.copy)(src, ref / name)
        (ref / "overlay.yaml").write_text(OVERLAY.format(prefix=Path(td) / "out" / "res"))
        (ref / "drive.py").write_text(DRIVER)
        git = ["git", "-c", "user.email=t@example.org", "-c", "user.name=t"]
        for cmd in (["init", "-q"], ["add", "-A"], ["commit", "-qm", "reference copy"]):
            subprocess.run([*git, *cmd], cwd=ref, check=True)
        out = HERE / "rng_diagnostics_fast42.json"
        subprocess.run([sys.executable, str(ref / "drive.py"), str(ref / "src"), str(ref / "overlay.yaml"),
                        str(out)], cwd=ref, check=True)
    data = json.loads(out.read_text())
    # the two constant text columns are stored once
    data["note"] = data["rows"][0]["note"]
    data["sequence_order"] = {r["summary_level"]: r["sequence_order"] for r in data["rows"]}
    for r in data["rows"]:
        del r["note"], r["sequence_order"]
    out.write_text(json.dumps(data, sort_keys=True, separators=(",", ":")) + "\n")
    print(out, len(data["rows"]), "rows;", data["summary"])


if __name__ == "__main__":
    main()
