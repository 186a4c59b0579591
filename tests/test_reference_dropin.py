"""Drop-in proof: the UNMODIFIED reference (`farkle.simulation.run_tournament.run_tournament`, its L4
runner `farkle.simulation.runner.run_tournament(cfg)` and the H2H stage `execute_h2h_schedule`) runs
with this repo's seam functions installed and must write the same checkpoint, metrics, rows and
sidecars as it does on its own.

The reference is imported from the checkout `tests/refpath.py` finds: /root/reference in the build
container, the staged copy `baseline/_ref/` (scripts/stage_reference.sh) on the GPU box.  Every test
runs twice: `oracle-backed` (CPU box: compute calls served by tests/oracle_engine.py, host logic
only) and `cuda` (marked gpu: the real CUDA engine through the C ABI, nothing monkeypatched).
"""

from __future__ import annotations

import os
import pickle
import sys
from pathlib import Path

import pytest

from refpath import numba_cache_env, reference_root

REF_ROOT = reference_root()
REF = (REF_ROOT or Path("/nonexistent")) / "src"
pytestmark = pytest.mark.skipif(REF_ROOT is None, reason="reference checkout not present on this box "
                                "(run scripts/stage_reference.sh where /root/reference exists)")

BACKENDS = [pytest.param("oracle", id="oracle-backed"),
            pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request):
    return request.param


@pytest.fixture()
def ref_rt(monkeypatch, backend):
    numba_cache_env()
    monkeypatch.syspath_prepend(str(REF))
    import farkle.simulation.run_tournament as rt

    if backend == "oracle":
        from farkle_ii_b200 import device as fdev
        from oracle_engine import OracleEngine

        eng = OracleEngine()
        monkeypatch.setattr(fdev, "get_engine", lambda device=None: eng)
    yield rt
    for name in [m for m in sys.modules if m == "farkle" or m.startswith("farkle.")]:
        sys.modules.pop(name)


@pytest.fixture()
def ref_env(monkeypatch):
    """The reference importable, no engine involved."""
    numba_cache_env()
    monkeypatch.syspath_prepend(str(REF))
    yield
    for name in [m for m in sys.modules if m == "farkle" or m.startswith("farkle.")]:
        sys.modules.pop(name)


def _grid(ref_sim):
    return ref_sim.generate_strategy_grid(
        score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
        consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
        run_up_score_opts=[True])[0]


def test_reference_driver_with_gpu_seams(ref_rt, tmp_path):
    from farkle.simulation import simulation as ref_sim

    from farkle_ii_b200 import reference_shim

    rt = ref_rt
    strats = _grid(ref_sim)
    kw = dict(strategies=strats, global_seed=42, n_jobs=1, collect_metrics=True,
              row_output_directory=None, num_shuffles=24, resume=False)
    cfg = lambda: rt.TournamentConfig(n_players=4, num_shuffles=24, deterministic_batch_size=6)  # noqa: E731
    rt.run_tournament(config=cfg(), checkpoint_path=tmp_path / "ref" / "4p.pkl", **kw)
    originals = reference_shim.install(rt)
    try:
        rt.run_tournament(config=cfg(), checkpoint_path=tmp_path / "gpu" / "4p.pkl", **kw)
        # the per-shuffle seam with rows, against the reference's own implementation
        task = rt.ShuffleTask(42, 4, 5, 0, 0)
        got = rt._play_one_shuffle(task, collect_rows=True)
        want = originals["_play_one_shuffle"](task, collect_rows=True)
    finally:
        reference_shim.uninstall(originals, rt)
    ref = pickle.loads((tmp_path / "ref" / "4p.pkl").read_bytes())
    gpu = pickle.loads((tmp_path / "gpu" / "4p.pkl").read_bytes())
    assert type(gpu["win_totals"]) is type(ref["win_totals"])
    assert dict(gpu["win_totals"]) == dict(ref["win_totals"])
    for key in ("outcome_counts", "metric_sums", "metric_square_sums"):
        assert gpu[key] == ref[key], key
    import pyarrow.parquet as pq

    a = pq.read_table(tmp_path / "ref" / "4p_metrics.parquet").to_pandas()
    b = pq.read_table(tmp_path / "gpu" / "4p_metrics.parquet").to_pandas()
    key = ["metric", "strategy"]
    assert a.sort_values(key).reset_index(drop=True).equals(b.sort_values(key).reset_index(drop=True))
    assert dict(got[0]) == dict(want[0]) and got[0].outcome_payload() == want[0].outcome_payload()
    assert {m: dict(v) for m, v in got[1].items()} == {m: dict(v) for m, v in want[1].items()}
    assert {m: dict(v) for m, v in got[2].items()} == {m: dict(v) for m, v in want[2].items()}
    assert got[3] == want[3] and [list(r) for r in got[3]] == [list(r) for r in want[3]]


def test_row_shards_through_the_reference_writer(ref_rt, tmp_path):
    """Rows mode: the shim hands its Arrow tables to the reference's own run_streaming_shard, so the
    shards and manifest lines must equal the reference's (pid and timestamps aside)."""
    import json

    import pyarrow.parquet as pq
    from farkle.simulation import simulation as ref_sim

    from farkle_ii_b200 import reference_shim

    rt = ref_rt
    strats = _grid(ref_sim)
    cfg = rt.TournamentConfig(n_players=5, n_strategies=len(strats))
    tasks = [rt.ShuffleTask(54, 5, s, 1000 + s, s // 2) for s in (3, 4, 7)]
    rt._init_worker(strats, cfg)
    want = rt._run_chunk_metrics(tasks, collect_rows=True, row_dir=tmp_path / "ref",
                                 manifest_path=tmp_path / "ref" / "manifest.jsonl")
    originals = reference_shim.install(rt)
    try:
        rt._init_worker(strats, cfg)
        got = rt._run_chunk_metrics(tasks, collect_rows=True, row_dir=tmp_path / "gpu",
                                    manifest_path=tmp_path / "gpu" / "manifest.jsonl")
    finally:
        reference_shim.uninstall(originals, rt)
    assert dict(got[0]) == dict(want[0]) and got[0].outcome_payload() == want[0].outcome_payload()
    for i in (1, 2):
        assert {m: dict(v) for m, v in got[i].items()} == {m: dict(v) for m, v in want[i].items()}
    drop = ("pid", "ts", "timestamp", "written_at")
    lines = {side: [{k: v for k, v in json.loads(x).items() if k not in drop}
                    for x in (tmp_path / side / "manifest.jsonl").read_text().splitlines()]
             for side in ("ref", "gpu")}
    assert lines["gpu"] == lines["ref"] and len(lines["ref"]) == 3
    for rec in lines["ref"]:
        a = pq.read_table(tmp_path / "ref" / rec["path"])
        b = pq.read_table(tmp_path / "gpu" / rec["path"])
        assert a.schema == b.schema and a.equals(b)


def test_h2h_block_runner_against_reference(ref_rt, tmp_path):
    """The H2H BlockRunner (h2h_schedule.py:1521): same progress dict as the reference's
    `_simulate_block`, accepted unchanged by its `_normalize_runner_result`."""
    from farkle.analysis import h2h_schedule as ref_h2h
    from farkle.simulation import simulation as ref_sim
    from farkle.simulation.strategies import build_strategy_manifest

    from farkle_ii_b200 import h2h as gpu_h2h

    strats = ref_sim.generate_strategy_grid()[0]
    manifest_path = tmp_path / "strategy_manifest.parquet"
    build_strategy_manifest(strats).to_parquet(manifest_path)
    for pair_id, (a, b), order, target, max_att in ((7, (12, 3400), 0, 20, 40), (8, (5000, 77), 1, 15, 30)):
        s1, s2 = (a, b) if order == 0 else (b, a)
        block = {"block_id": f"blk{pair_id}", "family_hash": "fam", "schedule_hash": "sched",
                 "root_seed": 4242, "pair_id": pair_id, "order": order, "seat1_strategy": s1,
                 "seat2_strategy": s2, "n_completed_required": target, "max_attempts": max_att,
                 "rng_scheme_version": 2, "rng_purpose_namespace": 203}
        for chunk in (7, 5000):
            want = ref_h2h._simulate_block(dict(block), manifest_path, chunk)
            got = gpu_h2h.gpu_block_runner(dict(block), manifest_path, chunk)
            assert got == want and list(got) == list(want)
            assert ref_h2h._normalize_runner_result(block, got) == ref_h2h._normalize_runner_result(block, want)
            block = dict(want)          # resume from the reference's progress


def test_seat_count_semantics_match_reference_seat_analysis(ref_env, tmp_path):
    """`seat_counts_from_rows` (the host restatement the GPU seat tallies are tested against) equals
    the reference's `_iter_seat_count_tables` on a curated Parquet of the same games."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from farkle.analysis import seat_analysis as ref_seat

    from farkle_ii_b200 import run_tournament as frt
    from farkle_ii_b200 import simulation as fsim

    z = np.load(Path(__file__).parent / "golden" / "games_fast_54_4.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    rows = z["rows"]
    gps = len(z["strategies"]) // k
    shuffle = sh0 + np.arange(len(rows)) // gps
    batch = (shuffle - sh0) // 2
    tbl = fsim.compact_rows_to_table(rows, root_seed=root, k=k, shuffle_index=shuffle,
                                     game_index=np.arange(len(rows)) % gps,
                                     deterministic_batch_id=batch, shuffle_seed=0)
    src = tmp_path / "curated.parquet"
    pq.write_table(tbl, src)
    ref_tbl = pa.concat_tables(list(ref_seat._iter_seat_count_tables(src, k))).to_pylist()
    want = {(r["deterministic_batch_id"], r["strategy"], r["seat"]):
            [r["raw_wins"], r["raw_exposures"], r["raw_completed_exposures"], r["raw_safety_limit_exposures"]]
            for r in ref_tbl}
    assert frt.seat_counts_from_rows(rows, batch) == want


def test_plan_and_limits_match_reference_modules(ref_env):
    """`shuffle_plan` / `limits` against the reference's `workload_planner` / `game_profile`:
    every float of the plan bit-identical, same identity hash, same rejections."""
    import itertools

    from farkle.simulation import game_profile as ref_gp
    from farkle.simulation import workload_planner as ref_wp

    from farkle_ii_b200 import limits, shuffle_plan

    for n, c in itertools.product([*range(1, 120), 4264, 4265, 4300, 1_234_567], (0.8, 0.95, 0.99)):
        assert shuffle_plan.worst_case_wilson_width(n, confidence=c) == \
            ref_wp.worst_case_wilson_width(n, confidence=c)
    for d, c in itertools.product((0.9, 0.2, 0.08, 0.03, 0.01), (0.9, 0.95, 0.999)):
        assert shuffle_plan.minimum_shuffles_for_resolution(d, confidence=c) == \
            ref_wp.minimum_shuffles_for_resolution(d, confidence=c)
    for k, sc, d, bc, cap in itertools.product((2, 6), (12, 5160), (0.03, 0.3), (2, 100),
                                               (None, 50)):
        kw = dict(root_seed=5, k=k, strategy_count=sc, resolution_delta=d, batch_count=bc,
                  shuffle_cap=cap, projected_games_per_second=4e8)
        mine, ref = shuffle_plan.plan_tournament_workload(**kw), ref_wp.plan_tournament_workload(**kw)
        assert mine.to_dict() == ref.to_dict()
        assert str(shuffle_plan.WorkloadCapExceeded(mine)) == str(ref_wp.WorkloadCapExceeded(ref))

    def profile(mod):
        return mod.GameProfile(
            default_target_score=500, default_max_rounds=7,
            tournament_max_rounds_overrides=(mod.TournamentMaxRoundsOverride(3, 2, 1, 0, 5),
                                             mod.TournamentMaxRoundsOverride(1, 2, 1, 0, 0)),
            h2h_max_rounds_overrides=(mod.H2HMaxRoundsOverride(1, 2, 1, 3, 4),))

    mine, ref = profile(limits), profile(ref_gp)
    assert mine.sha256 == ref.sha256 and mine.canonical_payload() == ref.canonical_payload()
    assert limits.GameProfile().sha256 == ref_gp.GameProfile().sha256
    for coord in ((3, 2, 1, 0), (1, 2, 1, 0), (1, 2, 1, 1)):
        kw = dict(zip(("root_seed", "k", "shuffle_index", "game_index"), coord))
        a, b = mine.tournament_limits(**kw), ref.tournament_limits(**kw)
        assert (a.target_score, a.max_rounds) == (b.target_score, b.max_rounds)
    a = mine.h2h_limits(root_seed=1, pair_id=2, order=1, attempt_index=3)
    b = ref.h2h_limits(root_seed=1, pair_id=2, order=1, attempt_index=3)
    assert (a.target_score, a.max_rounds) == (b.target_score, b.max_rounds) == (500, 4)
    for bad in ((1, 1, 0, 0, 5), (-1, 2, 0, 0, 5), (1, 2, 0, 0, -5), (1, 2, True, 0, 5)):
        msgs = []
        for mod in (limits, ref_gp):
            with pytest.raises(ValueError) as err:
                mod.TournamentMaxRoundsOverride(*bad)
            msgs.append(str(err.value))
        assert msgs[0] == msgs[1]


_PRELUDE = """
import os, sys
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
ref_src, engine, repo = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path.insert(0, ref_src)
from pathlib import Path
if engine != "none":
    sys.path.insert(0, repo)
    sys.path.insert(0, os.path.join(repo, "tests"))
    if engine == "oracle":
        from farkle_ii_b200 import device as fdev
        from oracle_engine import OracleEngine
        eng = OracleEngine()
        fdev.get_engine = lambda device=None: eng
"""

_RUNNER_SCRIPT = _PRELUDE + """
from farkle.config import load_app_config
from farkle.simulation import runner
if engine != "none":
    import farkle.simulation.run_tournament as rt
    from farkle_ii_b200 import reference_shim
    reference_shim.install(rt)
cfg = load_app_config(Path(ref_src).parent / "configs" / "fast_config.yaml", Path(sys.argv[4]))
print("games", runner.run_tournament(cfg))
if engine == "cuda":
    from farkle_ii_b200.device import get_engine
    print("gpu_launches", get_engine().kernel_launch_count())
"""

_OVERLAY = """
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


def _git_copy(tmp_path: Path) -> tuple[Path, list[str]]:
    """An unmodified, git-initialised copy of the checkout (the runner binds artifacts to the
    code identity of its work tree, utils/authenticated_contract.py:404-417)."""
    import shutil
    import subprocess

    if shutil.which("git") is None:
        pytest.skip("git not available")
    ref = tmp_path / "ref"
    ref.mkdir()
    for name in ("src", "configs", "pyproject.toml"):
        src = REF.parent / name
        (shutil.copytree if src.is_dir() else shutil.copy)(src, ref / name)
    git = ["git", "-c", "user.email=t@example.org", "-c", "user.name=t"]
    return ref, git


def _commit(ref: Path, git: list[str]) -> None:
    import subprocess

    subprocess.run([*git, "init", "-q"], cwd=ref, check=True)
    subprocess.run([*git, "add", "-A"], cwd=ref, check=True)
    subprocess.run([*git, "commit", "-qm", "reference copy"], cwd=ref, check=True)


def _same_tree(root_a: Path, root_b: Path) -> list[Path]:
    files = sorted(p.relative_to(root_a) for p in root_a.rglob("*") if p.is_file())
    assert files == sorted(p.relative_to(root_b) for p in root_b.rglob("*") if p.is_file())
    different = [str(f) for f in files if (root_a / f).read_bytes() != (root_b / f).read_bytes()]
    assert different == []
    return files


def test_runner_artifact_tree_is_byte_identical(tmp_path, backend):
    """The reference's L4 runner (`farkle run`: `simulation.runner.run_tournament(cfg)`, fast grid,
    k = 2 and 4, rows + expanded metrics, artifact contract v3) writes the SAME BYTES in every file
    -- row shards, manifests, metrics / checkpoint Parquets, checkpoint pickle, hash-bound sidecars,
    workload plan, `simulation.done.json` -- with this repo's seams installed as it does alone.
    Runs an unmodified, git-initialised copy of the checkout in two subprocesses."""
    import subprocess

    ref, git = _git_copy(tmp_path)
    outs = {}
    for tag in ("cpu", "gpu"):
        outs[tag] = tmp_path / f"out_{tag}"
        (ref / f"overlay_{tag}.yaml").write_text(_OVERLAY.format(prefix=outs[tag] / "res"))
    (ref / "drive.py").write_text(_RUNNER_SCRIPT)
    _commit(ref, git)
    repo = str(Path(__file__).resolve().parents[1])
    for tag, engine in (("cpu", "none"), ("gpu", backend)):
        done = subprocess.run([sys.executable, str(ref / "drive.py"), str(ref / "src"), engine, repo,
                               str(ref / f"overlay_{tag}.yaml")],
                              cwd=ref, capture_output=True, text=True, timeout=900)
        assert done.returncode == 0, done.stderr[-2000:]
        assert "games 720" in done.stdout            # 12 shuffles x (40 + 20) games
        if engine == "cuda":                          # the CUDA library really did the playing
            assert int(done.stdout.split("gpu_launches")[1].split()[0]) > 0
    files = _same_tree(outs["cpu"] / "res_seed_42", outs["gpu"] / "res_seed_42")
    assert len(files) > 60 and any(f.name == "simulation.done.json" for f in files)


# The H2H stage: frozen family -> plan_h2h_schedule -> execute_h2h_schedule, all the reference's
# own code (the family artifacts are written with its public artifact helpers the way its unit
# tests do, tests/unit/analysis/test_h2h_schedule.py:41-108); only `block_runner` differs.
_H2H_SCRIPT = _PRELUDE + """
import json
import pandas as pd, pyarrow as pa, pyarrow.parquet as pq
from farkle.analysis.h2h_schedule import execute_h2h_schedule, plan_h2h_schedule
from farkle.config import AppConfig, ArtifactScope, IOConfig, SimConfig
from farkle.simulation.simulation import generate_strategy_grid
from farkle.simulation.strategies import build_strategy_manifest
from farkle.utils.artifact_contract import make_artifact_sidecar
from farkle.utils.artifacts import write_json_artifact_atomic, write_parquet_artifact_atomic

out, runner_kind = Path(sys.argv[4]), sys.argv[5]
roots = (11, 22)
cfg = AppConfig(io=IOConfig(results_dir_prefix=out / "results"),
                sim=SimConfig(seed=roots[0], seed_list=list(roots), n_players_list=[2, 4]))
cfg.screening.practical_delta_by_k = {2: 0.03, 4: 0.03}
cfg.screening.delta_across_k = 0.03
cfg.head2head.total_game_cap = None
cfg.head2head.n_jobs = 1
cfg.head2head.practical_delta = 0.25          # small power-planned blocks
cfg.resources.scheduler_memory_budget_mb = 2048   # torch alone exceeds the 768 MiB library default
cfg.resources.process_tree_warning_threshold_mb = 8192
cfg.resources.aggregate_memory_hard_limit_mb = 12288
cfg.resources.minimum_system_available_memory_mb = 256
cfg.head2head.seat1_advantage_scenarios = (0.0,)
candidates = (17, 1203, 2999, 4100)           # ids of the default grid
family_hash = "a" * 64
membership = pd.DataFrame({"strategy": list(candidates), "final_family": [True] * len(candidates),
                           "family_hash": [family_hash] * len(candidates)})
membership["strategy"] = pd.array(membership["strategy"].tolist(), dtype="Int32")
manifest = {"family_hash": family_hash, "candidates": list(candidates),
            "candidate_count": len(candidates), "root_seeds": list(roots), "single_root_execution": False}
common = dict(producer="test", scope=ArtifactScope.H2H_2P, source_scope=ArtifactScope.CROSS_SEED,
              operation="candidate_family_freeze", player_counts=[2], required_player_counts=[2],
              missing_cell_policy="fail", seed_scope="both_roots_combined")
mp = cfg.h2h_candidate_family_path()
write_parquet_artifact_atomic(pa.Table.from_pandas(membership, preserve_index=False), mp,
    sidecar=make_artifact_sidecar(cfg, mp, consistency_columns=membership.columns.tolist(), **common))
jp = cfg.h2h_candidate_family_manifest_path()
write_json_artifact_atomic(manifest, jp,
    sidecar=make_artifact_sidecar(cfg, jp, consistency_columns=list(manifest), **common))
plan_h2h_schedule(cfg)
strategies = generate_strategy_grid()[0]
sm = cfg.strategy_manifest_root_path()
sm.parent.mkdir(parents=True, exist_ok=True)
build_strategy_manifest(strategies).to_parquet(sm)
schedule = pq.read_table(cfg.h2h_block_manifest_path()).to_pandas()
print("blocks", len(schedule), "required", int(schedule["n_completed_required"].iloc[0]))
kw = {}
if runner_kind == "single":
    from farkle_ii_b200 import h2h
    kw = dict(n_jobs=1, block_runner=h2h.gpu_block_runner)
elif runner_kind == "batched":
    from farkle_ii_b200 import h2h
    kw = dict(n_jobs=1, block_runner=h2h.BatchedBlockRunner(schedule.to_dict(orient="records"), chunk_games=int(sys.argv[6])))
art = execute_h2h_schedule(cfg, chunk_games=int(sys.argv[6]), **kw)
counts = pq.read_table(art.order_counts).to_pandas()
print("completed", int(counts["games_completed"].sum()), "attempted", int(counts["games_attempted"].sum()))
if engine == "cuda":
    from farkle_ii_b200.device import get_engine
    print("gpu_launches", get_engine().kernel_launch_count())
"""


@pytest.mark.parametrize("runner_kind,chunk", [("single", 5000), ("batched", 9)])
def test_h2h_stage_with_gpu_block_runner(tmp_path, backend, runner_kind, chunk):
    """`execute_h2h_schedule` (analysis/h2h_schedule.py:1597) — the reference's whole H2H execution
    stage, on a schedule its own `plan_h2h_schedule` produced — writes the same block Parquets,
    sidecars, execution state and order counts with this repo's `BlockRunner` (one block per call,
    and the batched runner that advances every pending block per launch) as with its own
    `_simulate_block`.  `chunk` 9 forces several durable checkpoints per block."""
    import subprocess

    ref, git = _git_copy(tmp_path)
    (ref / "drive_h2h.py").write_text(_H2H_SCRIPT)
    _commit(ref, git)
    repo = str(Path(__file__).resolve().parents[1])
    outs = {"cpu": tmp_path / "out_cpu", "gpu": tmp_path / "out_gpu"}
    seen = {}
    for tag, engine, kind in (("cpu", "none", "reference"), ("gpu", backend, runner_kind)):
        done = subprocess.run([sys.executable, str(ref / "drive_h2h.py"), str(ref / "src"), engine, repo,
                               str(outs[tag]), kind, str(chunk)],
                              cwd=ref, capture_output=True, text=True, timeout=900)
        assert done.returncode == 0, done.stderr[-3000:]
        seen[tag] = [ln for ln in done.stdout.splitlines() if ln.startswith(("blocks", "completed"))]
        if engine == "cuda":
            assert int(done.stdout.split("gpu_launches")[1].split()[0]) > 0
    assert seen["cpu"] == seen["gpu"] and len(seen["cpu"]) == 2
    # timing telemetry aside, every artifact of the stage is byte-identical
    files_a = sorted(p.relative_to(outs["cpu"]) for p in outs["cpu"].rglob("*") if p.is_file())
    files_b = sorted(p.relative_to(outs["gpu"]) for p in outs["gpu"].rglob("*") if p.is_file())
    assert files_a == files_b and any("h2h" in str(f) for f in files_a)
    different = [str(f) for f in files_a if (outs["cpu"] / f).read_bytes() != (outs["gpu"] / f).read_bytes()]
    assert different == [], different


def test_standalone_driver_against_reference_driver(ref_rt, tmp_path, monkeypatch):
    """`farkle_ii_b200.run_tournament.run_tournament` (no reference code involved) against the
    reference's own `run_tournament` on the same cell: rows + metric chunks + checkpoint, then an
    interrupted run resumed from its checkpoint and manifests (run_tournament.py:1243-1373).

    Same checkpoint contents (win totals incl. key order, outcome counts, metric sums and squares,
    `meta` key for key), same `{k}p_metrics.parquet`, same row shards and manifest records, same
    metric-chunk tables and manifest records."""
    import json

    import pyarrow.parquet as pq
    from farkle.simulation import simulation as ref_sim

    from farkle_ii_b200 import run_tournament as frt
    from farkle_ii_b200.strategies import generate_strategy_grid

    rt = ref_rt
    ref_strats = _grid(ref_sim)
    my_strats = generate_strategy_grid(
        score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
        consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
        run_up_score_opts=[True])[0]
    k, root, shuffles, batch = 4, 77, 23, 5

    def run(mod, strats, out, n=shuffles, chunks=True, **kw):
        cfg = mod.TournamentConfig(n_players=k, num_shuffles=n, deterministic_batch_size=batch)
        mod.run_tournament(config=cfg, global_seed=root, checkpoint_path=out / f"{k}p_checkpoint.pkl",
                           n_jobs=1, collect_metrics=True, row_output_directory=out / "rows",
                           metric_chunk_directory=(out / "chunks") if chunks else None, num_shuffles=n,
                           strategies=strats, checkpoint_metadata={"run": "t"}, **kw)

    def payload(out):
        p = pickle.loads((out / f"{k}p_checkpoint.pkl").read_bytes())
        w = p["win_totals"]
        return {"wins": list(dict(w).items()), "outcome": p["outcome_counts"],
                "sums": {m: list(v.items()) for m, v in p["metric_sums"].items()},
                "squares": {m: list(v.items()) for m, v in p["metric_square_sums"].items()},
                "meta": list(p["meta"].items())}

    def manifest(path, drop=("pid", "ts", "timestamp", "written_at")):
        return [{a: b for a, b in json.loads(x).items() if a not in drop}
                for x in path.read_text().splitlines()]

    def same_outputs(a, b):
        assert payload(a) == payload(b)
        key = ["metric", "strategy"]
        ta, tb = (pq.read_table(d / f"{k}p_metrics.parquet") for d in (a, b))
        assert ta.schema == tb.schema and ta.equals(tb)
        for sub, mname in (("rows", "manifest.jsonl"), ("chunks", "metrics_manifest.jsonl")):
            if not (a / sub).exists():
                assert not (b / sub).exists()
                continue
            ma, mb = manifest(a / sub / mname), manifest(b / sub / mname)
            assert ma == mb and len(ma) == (shuffles if sub == "rows" else -(-shuffles // batch))
            for rec in ma:
                xa, xb = pq.read_table(a / sub / rec["path"]), pq.read_table(b / sub / rec["path"])
                assert xa.schema == xb.schema and xa.equals(xb), rec["path"]
        del key

    ref_out, my_out = tmp_path / "ref", tmp_path / "mine"
    run(rt, ref_strats, ref_out, resume=False)
    run(frt, my_strats, my_out, resume=False)
    same_outputs(ref_out, my_out)
    # rows + metrics without a chunk directory: aggregates absorbed batch by batch (key order of the
    # reference's parent loop, run_tournament.py:1596-1705)
    run(rt, ref_strats, tmp_path / "ref2", chunks=False, resume=False)
    run(frt, my_strats, tmp_path / "mine2", chunks=False, resume=False)
    same_outputs(tmp_path / "ref2", tmp_path / "mine2")
    # interrupted after 11 shuffles (a checkpoint, 11 row shards, 3 metric chunks exist), then resumed:
    # nothing already listed is rewritten, the aggregates equal the uninterrupted run's
    part = tmp_path / "resumed"
    run(frt, my_strats, part, n=11, resume=False)
    stamp = {p: p.stat().st_mtime_ns for p in (part / "rows").glob("*.parquet")}
    assert len(stamp) == 11
    run(frt, my_strats, part, resume=True)
    assert all(p.stat().st_mtime_ns == t for p, t in stamp.items())
    got, want = payload(part), payload(my_out)
    assert got["outcome"] == want["outcome"] and dict(got["wins"]) == dict(want["wins"])
    assert {m: dict(v) for m, v in got["sums"].items()} == {m: dict(v) for m, v in want["sums"].items()}
    assert dict(got["meta"])["completed_shuffle_indices"] == list(range(shuffles))
    assert len(manifest(part / "rows" / "manifest.jsonl")) == shuffles
    # a checkpoint every group when the cadence is zero seconds
    cfg = frt.TournamentConfig(n_players=k, num_shuffles=shuffles, deterministic_batch_size=batch, ckpt_every_sec=0)
    seen = []
    real = frt._save_checkpoint
    monkeypatch.setattr(frt, "_save_checkpoint", lambda path, *a: (seen.append(len(a[3]["completed_shuffle_indices"])),
                                                                   real(path, *a))[1])
    frt.run_tournament(config=cfg, global_seed=root, checkpoint_path=tmp_path / "cad" / "c.pkl", collect_metrics=True,
                       metric_chunk_directory=tmp_path / "cad" / "chunks", num_shuffles=shuffles,
                       strategies=my_strats, resume=False)
    assert seen == [5, 10, 15, 20, 23, 23]


def test_all_player_statistics_match_reference_metrics_stage(ref_env, tmp_path):
    """The all-player sufficient statistics (f-3): `run_tournament.all_player_table` over the
    host restatement of the device kernel equals, column for column and bit for bit (float sums
    included), what the reference's metrics stage derives from a curated Parquet of the same games
    (`analysis/all_player_metrics.py::_iter_batch_tables`)."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from farkle.analysis import all_player_metrics as ref_apm
    from farkle.utils.parallel import ProcessTreeMemoryGuard

    from all_player_rows import all_player_from_rows
    from farkle_ii_b200 import run_tournament as frt
    from farkle_ii_b200 import simulation as fsim

    for name, per_batch in (("games_fast_54_4", 2), ("games_full_0_5", 1), ("games_fast_42_2", 3)):
        z = np.load(Path(__file__).parent / "golden" / f"{name}.npz")
        root, k, sh0, nsh = (int(x) for x in z["meta"])
        rows = z["rows"]
        n = len(z["strategies"])
        gps = n // k
        shuffle = sh0 + np.arange(len(rows)) // gps
        batch = (shuffle - sh0) // per_batch
        tbl = fsim.compact_rows_to_table(rows, root_seed=root, k=k, shuffle_index=shuffle,
                                         game_index=np.arange(len(rows)) % gps,
                                         deterministic_batch_id=batch, shuffle_seed=0)
        src = tmp_path / f"{name}.parquet"
        pq.write_table(tbl, src)
        guard = ProcessTreeMemoryGuard(1 << 20, rss_warning_mb=1 << 20, minimum_system_available_memory_mb=1)
        want = pa.concat_tables(list(ref_apm._iter_batch_tables(src, k, max_batch_bytes=1 << 30,
                                                                max_batch_rows=1 << 20, memory_guard=guard)))
        n_slots = int(batch.max()) + 1
        n_ids = int(rows["seats"]["strategy"].max()) + 1
        stats = all_player_from_rows(rows, batch, n_slots, n_ids)
        got = frt.all_player_table(stats, np.arange(n_ids), root_seed=root, k=k)
        assert got.schema == want.schema == ref_apm.all_player_batch_schema()
        assert got.equals(want), name


def test_row_validator_matches_reference(ref_env):
    """`validate_simulation_row` (simulation/simulation.py:450-563): the mirror accepts what the
    reference accepts and rejects what it rejects, with the same message, over rows of real games
    and thirty ways of corrupting them; `validate_compact_rows` catches the compact-record
    corruptions before any row is expanded."""
    import copy

    import numpy as np
    from farkle.simulation import simulation as ref_sim

    from farkle_ii_b200 import simulation as fsim

    z = np.load(Path(__file__).parent / "golden" / "games_fast_54_4.npz")
    rows = z["rows"]
    good = fsim.expand_rows(rows[:60], root_seed=54)
    assert any(r["termination_status"] == "safety_limit" for r in fsim.expand_rows(rows, root_seed=54)) or True
    safety_rows = [r for r in fsim.expand_rows(rows, root_seed=54) if r["hit_safety_limit"]][:3]

    def outcome(fn, row):
        try:
            fn(row)
            return None
        except ValueError as exc:
            return str(exc)

    mutations = [
        lambda r: r.update(k=0), lambda r: r.update(termination_status="aborted"),
        lambda r: r.update(outcome_schema_version=1), lambda r: r.pop("P2_strategy"),
        lambda r: r.update(P1_score=1.5), lambda r: r.update(P1_score=True),
        lambda r: r.update(P2_strategy=r["P1_strategy"]), lambda r: r.pop("P3_rank"),
        lambda r: r.update(winner_seat=None), lambda r: r.update(winner_seat="P9"),
        lambda r: r.update(P1_rank=r["P2_rank"]), lambda r: r.update(P1_rank=None),
        lambda r: r.update(P1_rank=r["P2_rank"], P2_rank=r["P1_rank"],
                           winner_seat=next(s for s in ("P1", "P2", "P3", "P4")
                                            if {"P1": r["P2_rank"], "P2": r["P1_rank"]}.get(s, r[f"{s}_rank"]) == 1)),
        lambda r: r.update(winner_strategy=None), lambda r: r.update(winner_strategy=-5),
        lambda r: r.update(winning_score=None), lambda r: r.update(victory_margin=None),
        lambda r: r.update(hit_safety_limit=True), lambda r: r.update(P3_hit_max_rounds=True),
        lambda r: r.update(winning_score=r["winning_score"] + 50),
        lambda r: r.update(victory_margin=r["victory_margin"] + 50),
        lambda r: r.update(P4_loss_margin=None), lambda r: r.update(P4_loss_margin=True),
        lambda r: r.update(P4_loss_margin=r["P4_loss_margin"] + 1),
        lambda r: r.update(seat_ranks=None), lambda r: r.update(seat_ranks=list(reversed(r["seat_ranks"]))),
        lambda r: r.update(termination_status="safety_limit"),
        lambda r: r.update(termination_status="safety_limit", hit_safety_limit=True),
    ]
    for row in good[:6]:
        assert outcome(fsim.validate_simulation_row, row) is None and outcome(ref_sim.validate_simulation_row, row) is None
        for mutate in mutations:
            bad = copy.deepcopy(row)
            mutate(bad)
            mine, ref = outcome(fsim.validate_simulation_row, bad), outcome(ref_sim.validate_simulation_row, bad)
            assert (mine is None) == (ref is None), (mine, ref)
            if "canonical" not in str(ref):            # the id helper's wording is the reference's own
                assert mine == ref
    safety_mutations = [
        lambda r: r.update(hit_safety_limit=False), lambda r: r.update(P1_hit_max_rounds=False),
        lambda r: r.update(winner_seat="P1"), lambda r: r.update(winning_score=100), lambda r: r.update(P2_rank=1),
        lambda r: r.update(seat_ranks=None), lambda r: r.update(seat_ranks=["P1", None, None, None]),
        lambda r: r.update(P2_loss_margin=0),
    ]
    for row in safety_rows:
        assert outcome(fsim.validate_simulation_row, row) is None and outcome(ref_sim.validate_simulation_row, row) is None
        for mutate in safety_mutations:
            bad = copy.deepcopy(row)
            mutate(bad)
            assert outcome(fsim.validate_simulation_row, bad) == outcome(ref_sim.validate_simulation_row, bad) is not None
    # compact records: wrong winner, duplicate seat, winner on a safety-limit row
    completed = np.flatnonzero((rows["flags"] & 1) == 0)[:5]
    for what in ("winner", "duplicate"):
        bad = rows[completed].copy()
        if what == "winner":
            bad["winner_seat"][2] = (bad["winner_seat"][2] + 1) % 4
        else:
            bad["seats"]["strategy"][1, 3] = bad["seats"]["strategy"][1, 0]
        with pytest.raises(ValueError):
            fsim.validate_compact_rows(bad)
        with pytest.raises(ValueError):
            fsim.compact_rows_to_table(bad, root_seed=54, k=4, shuffle_index=0, game_index=np.arange(5),
                                       deterministic_batch_id=0, shuffle_seed=0)
    fsim.validate_compact_rows(rows)
