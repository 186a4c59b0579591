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
