"""CPU tests of the host mirror: strategy grid order/ids, packing, library exports."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from farkle_ii_b200 import _native, layout
from farkle_ii_b200.strategies import (
    FavorDiceOrScore,
    ThresholdStrategy,
    build_stop_at_strategy,
    generate_strategy_grid,
    pack_strategies,
    parse_strategy,
    prepare_strategy_ids,
    unpack_strategy,
)

FAST = dict(score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
            consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
            run_up_score_opts=[True])
TINY = dict(score_thresholds=[500], dice_thresholds=[2], smart_five_opts=[False],
            smart_one_opts=[False], consider_score_opts=[True], consider_dice_opts=[True],
            auto_hot_dice_opts=[False, True], run_up_score_opts=[False])


@pytest.mark.parametrize("fixture,kwargs,count", [
    ("games_fast_42_2.npz", FAST, 80), ("games_full_0_2.npz", {}, 5160), ("oracle12.npz", TINY, 4)])
def test_grid_matches_reference_table(golden_dir, fixture, kwargs, count):
    # the fixture's table was packed from the reference's generate_strategy_grid()
    want = np.load(golden_dir / fixture)["strategies"]
    strategies, meta = generate_strategy_grid(**kwargs)
    assert len(strategies) == count == len(meta)
    assert [s.strategy_id for s in strategies] == list(range(count))
    got = pack_strategies(strategies)
    assert got.dtype == layout.STRATEGY_DTYPE
    assert got.tobytes() == want.tobytes()
    assert list(meta["strategy_idx"]) == list(range(count))


def test_default_grid_inactive_sentinels():
    strategies, _ = generate_strategy_grid()
    assert {s.score_threshold for s in strategies if not s.consider_score} == {199}
    assert {s.dice_threshold for s in strategies if not s.consider_dice} == {-1}
    assert all(s.favor_dice_or_score is FavorDiceOrScore.DICE
               for s in strategies if s.consider_dice and not s.consider_score)


def test_stop_at_strategies_appended():
    base, _ = generate_strategy_grid(**FAST)
    both, meta = generate_strategy_grid(**FAST, include_stop_at=True,
                                        include_stop_at_heuristic=True)
    assert len(both) == len(base) + 8
    assert [str(s) for s in both[-8:]] == [
        *(f"stop_at_{t}" for t in (350, 400, 450, 500)),
        *(f"stop_at_{t}_heuristic" for t in (350, 400, 450, 500))]
    assert len(set(meta["strategy_id"])) == len(both)
    with pytest.raises(ValueError):
        build_stop_at_strategy(123)


def test_strategy_validation_and_roundtrip():
    with pytest.raises(ValueError):
        ThresholdStrategy(smart_one=True, smart_five=False)
    with pytest.raises(ValueError):
        ThresholdStrategy(require_both=True, consider_dice=False)
    s = ThresholdStrategy(score_threshold=350, dice_threshold=1, smart_five=True, smart_one=True,
                          require_both=True, auto_hot_dice=True,
                          favor_dice_or_score=FavorDiceOrScore.DICE)
    assert str(s) == "Strat(350,1)[SD][FOFD][AND][H-]"
    assert parse_strategy(str(s)) == s
    assert unpack_strategy(pack_strategies([s])[0]) == s
    with pytest.raises(ValueError):
        parse_strategy("Strat(1,2)")


def test_decide_truth_table():
    # reference tests/unit/simulation/test_strategies.py style: entry gate, final round, AND/OR
    s = ThresholdStrategy(score_threshold=300, dice_threshold=2)
    assert s.decide(turn_score=100, dice_left=1, has_scored=False)          # 500 entry gate
    assert not s.decide(turn_score=600, dice_left=1, has_scored=False)
    assert s.decide(turn_score=100, dice_left=4, has_scored=True)           # both unmet
    assert not s.decide(turn_score=300, dice_left=4, has_scored=True)       # OR: score hit
    both = ThresholdStrategy(score_threshold=300, dice_threshold=2, require_both=True)
    assert both.decide(turn_score=300, dice_left=4, has_scored=True)        # AND: dice unmet
    assert s.decide(turn_score=900, dice_left=1, has_scored=True, final_round=True,
                    score_to_beat=10_000, running_total=9_000)              # must catch up
    assert not s.decide(turn_score=100, dice_left=6, has_scored=True, final_round=True,
                        score_to_beat=10_000, running_total=10_050)         # ahead, no run-up


def test_prepare_strategy_ids():
    a, b, c = ThresholdStrategy(), ThresholdStrategy(strategy_id=0), ThresholdStrategy()
    assert prepare_strategy_ids([a, b, c]) == [1, 0, 2]
    with pytest.raises(ValueError):
        prepare_strategy_ids([ThresholdStrategy(strategy_id=3), ThresholdStrategy(strategy_id=3)])


def test_row_layout_matches_header():
    for k in range(1, 13):
        assert layout.row_dtype(k).itemsize == layout.row_stride(k)
        assert layout.row_stride(k) % 16 == 0 and layout.row_stride(k) >= 16 + 28 * k
    assert layout.row_stride(2) == 80 and layout.row_stride(12) == 352


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU box and exports what include/farkle_b200.h declares."""
    _native.build()
    lib = _native.lib()
    header = _native.HEADER.read_text()
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fb_abi_version() == 4
    for k in (2, 5, 12):
        assert lib.fb_row_stride(k) == layout.row_stride(k)
    assert lib.fb_workspace_bytes(2, 1000) >= 1000 * 2 * 36


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    lib = _native.lib()
    assert lib.fb_init(0) == -1  # FB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.fb_last_error()
    out = (ctypes.c_uint64 * 4)()
    assert lib.fb_seed_streams(None, 1, out, None) == -1
    from farkle_ii_b200.device import get_engine

    with pytest.raises(_native.NativeError):
        get_engine()


def test_workload_planner_kats():
    """The reference's own planner KATs (tests/unit/simulation/test_workload_planner.py:18-60) and
    the grid sizes bench.py uses."""
    from farkle_ii_b200.shuffle_plan import (minimum_shuffles_for_resolution,
                                                 plan_tournament_workload, worst_case_wilson_width)

    plan = plan_tournament_workload(root_seed=17, k=4, strategy_count=200, resolution_delta=0.03)
    assert plan.required_shuffles_unrounded == 4265
    assert worst_case_wilson_width(4264) > 0.03 and worst_case_wilson_width(4265) <= 0.03
    assert (plan.required_shuffles, plan.batch_count, plan.shuffles_per_batch) == (4300, 100, 43)
    assert plan.required_games == 215_000 and plan.achieved_resolution <= 0.03
    assert plan.batch_construction == "equal_contiguous" and plan.status == "not_started"
    low = plan_tournament_workload(root_seed=1, k=2, strategy_count=20, resolution_delta=0.9)
    assert minimum_shuffles_for_resolution(0.9) == 1
    assert (low.required_shuffles, low.shuffles_per_batch) == (3_000, 30)
    capped = plan_tournament_workload(root_seed=9, k=2, strategy_count=10, resolution_delta=0.03,
                                      shuffle_cap=4_000)
    assert capped.cap_exceeded and capped.status == "blocked_by_cap"
    assert capped.achieved_resolution_at_cap > 0.03
    fast = plan_tournament_workload(root_seed=42, k=2, strategy_count=80, resolution_delta=0.08,
                                    batch_count=20)
    assert (fast.required_shuffles, fast.shuffles_per_batch, fast.required_games) == (600, 30, 24_000)
    full = plan_tournament_workload(root_seed=42, k=2, strategy_count=5160, resolution_delta=0.03)
    assert full.required_games == 11_094_000
    assert full.with_games_per_second(4e8).projected_runtime_seconds == pytest.approx(0.0277, rel=1e-2)
    for bad in (dict(k=1), dict(strategy_count=7), dict(batch_count=1), dict(resolution_delta=1.5)):
        with pytest.raises(ValueError):
            plan_tournament_workload(**{**dict(root_seed=0, k=2, strategy_count=10, resolution_delta=0.1), **bad})


def test_lag_request_struct_layout_matches_header(tmp_path):
    """`_native.LagRequest` (ctypes) against `fb_lag_request_t` as a C compiler lays it out from
    include/farkle_b200.h: same size, same offset for every field."""
    import ctypes
    import shutil
    import subprocess

    from farkle_ii_b200 import _native

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    fields = [name for name, _ in _native.LagRequest._fields_]
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "farkle_b200.h"\nint main(void) {\n'
        '  printf("%zu\\n", sizeof(fb_lag_request_t));\n'
        + "".join(f'  printf("%zu\\n", offsetof(fb_lag_request_t, {f}));\n' for f in fields)
        + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    include = Path(__file__).resolve().parents[1] / "include"
    subprocess.run([gcc, "-std=c11", f"-I{include}", str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_native.LagRequest)
    assert out[1:] == [getattr(_native.LagRequest, f).offset for f in fields]
    assert ctypes.sizeof(_native.LagRequest) == 104          # ABI 4: + all_player_dev
    # fb_cell_t (fb_play_tournament_cells)
    cell_fields = [name for name, _ in _native.Cell._fields_]
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "farkle_b200.h"\nint main(void) {\n'
        '  printf("%zu\\n", sizeof(fb_cell_t));\n'
        + "".join(f'  printf("%zu\\n", offsetof(fb_cell_t, {f}));\n' for f in cell_fields)
        + "  return 0;\n}\n")
    subprocess.run([gcc, "-std=c11", f"-I{include}", str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_native.Cell) == 40
    assert out[1:] == [getattr(_native.Cell, f).offset for f in cell_fields]


def test_score_lookup_table_layout(tmp_path):
    """The triangular score table of csrc/scoring.cuh (924 entries per variant instead of 84 x 84):
    rowA/colB address every multiset of at most six dice exactly once and every entry carries the
    fields of `host_evaluate_counts` (scoring_lookup.py:123-172) -- checked on the host, no GPU."""
    import shutil
    import subprocess

    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "lut_check"
    subprocess.run([gxx, "-std=c++17", "-O1", "-o", str(exe), str(Path(__file__).parent / "lut_check.cpp")],
                   check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [924, 0, 924, 28496]
