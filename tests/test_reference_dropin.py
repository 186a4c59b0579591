"""Drop-in proof: the UNMODIFIED reference driver (`farkle.simulation.run_tournament.run_tournament`,
imported from /root/reference) runs with this repo's seam functions installed and must write the
same checkpoint, metrics and rows as it does on its own.

Only runs where /root/reference exists (the build container); compute calls are served by the
oracle-backed engine here — GPU parity of the same calls is covered by tests/test_gpu_parity.py and
tests/test_host_surface.py[cuda].
"""

from __future__ import annotations

import os
import pickle
import sys
from pathlib import Path

import pytest

REF = Path("/root/reference/src")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")


@pytest.fixture()
def ref_rt(monkeypatch):
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
    monkeypatch.syspath_prepend(str(REF))
    import farkle.simulation.run_tournament as rt

    from farkle_ii_b200 import device as fdev
    from oracle_engine import OracleEngine

    eng = OracleEngine()
    monkeypatch.setattr(fdev, "get_engine", lambda device=None: eng)
    yield rt
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


def test_seat_count_semantics_match_reference_seat_analysis(ref_rt, tmp_path):
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


def test_plan_and_limits_match_reference_modules(ref_rt):
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


_RUNNER_SCRIPT = """
import os, sys
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
ref_src, overlay, use_shim, repo = sys.argv[1], sys.argv[2], sys.argv[3] == "1", sys.argv[4]
sys.path.insert(0, ref_src)
from pathlib import Path
from farkle.config import load_app_config
from farkle.simulation import runner
if use_shim:
    sys.path.insert(0, repo)
    sys.path.insert(0, os.path.join(repo, "tests"))
    import farkle.simulation.run_tournament as rt
    from farkle_ii_b200 import device as fdev, reference_shim
    from oracle_engine import OracleEngine
    eng = OracleEngine()
    fdev.get_engine = lambda device=None: eng
    reference_shim.install(rt)
cfg = load_app_config(Path(ref_src).parent / "configs" / "fast_config.yaml", Path(overlay))
print("games", runner.run_tournament(cfg))
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


def test_runner_artifact_tree_is_byte_identical(tmp_path):
    """The reference's L4 runner (`farkle run`: `simulation.runner.run_tournament(cfg)`, fast grid,
    k = 2 and 4, rows + expanded metrics, artifact contract v3) writes the SAME BYTES in every file
    -- row shards, manifests, metrics / checkpoint Parquets, checkpoint pickle, hash-bound sidecars,
    workload plan, `simulation.done.json` -- with this repo's seams installed as it does alone.
    Runs an unmodified, git-initialised copy of the checkout in two subprocesses."""
    import shutil
    import subprocess

    if shutil.which("git") is None:
        pytest.skip("git not available")
    ref = tmp_path / "ref"
    ref.mkdir()
    for name in ("src", "configs", "pyproject.toml"):
        src = REF.parent / name
        (shutil.copytree if src.is_dir() else shutil.copy)(src, ref / name)
    outs = {}
    for tag in ("cpu", "gpu"):
        outs[tag] = tmp_path / f"out_{tag}"
        (ref / f"overlay_{tag}.yaml").write_text(_OVERLAY.format(prefix=outs[tag] / "res"))
    (ref / "drive.py").write_text(_RUNNER_SCRIPT)
    git = ["git", "-c", "user.email=t@example.org", "-c", "user.name=t"]
    subprocess.run([*git, "init", "-q"], cwd=ref, check=True)
    subprocess.run([*git, "add", "-A"], cwd=ref, check=True)
    subprocess.run([*git, "commit", "-qm", "reference copy"], cwd=ref, check=True)
    repo = str(Path(__file__).resolve().parents[1])
    for tag, shim in (("cpu", "0"), ("gpu", "1")):
        done = subprocess.run([sys.executable, str(ref / "drive.py"), str(ref / "src"),
                               str(ref / f"overlay_{tag}.yaml"), shim, repo],
                              cwd=ref, capture_output=True, text=True, timeout=600)
        assert done.returncode == 0, done.stderr[-2000:]
        assert "games 720" in done.stdout            # 12 shuffles x (40 + 20) games
    root_a, root_b = outs["cpu"] / "res_seed_42", outs["gpu"] / "res_seed_42"
    files = sorted(p.relative_to(root_a) for p in root_a.rglob("*") if p.is_file())
    assert files == sorted(p.relative_to(root_b) for p in root_b.rglob("*") if p.is_file())
    assert len(files) > 60 and any(f.name == "simulation.done.json" for f in files)
    different = [str(f) for f in files if (root_a / f).read_bytes() != (root_b / f).read_bytes()]
    assert different == []
