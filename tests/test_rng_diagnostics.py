"""RNG lag diagnostics of the strategy groups.

CPU: the host algebra (joining launches, the report rows) against brute force and -- where the
reference checkout exists -- against the reference's own ``_OnlineMetric`` / ``_rows_for_online_group``.
GPU: ``lag_gather_kernel`` / ``lag_edges_kernel`` against the same brute force over the oracle's rows.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

import oracle as fo
from farkle_ii_b200 import rng_diagnostics as rd
from farkle_ii_b200.layout import LAG_WIDTH
from oracle_engine import OracleEngine

GOLDEN = Path(__file__).parent / "golden"
REF = Path("/root/reference/src")


def _random_obs(rng, n, m, max_rounds=200):
    return (rng.integers(1, max_rounds + 1, size=(n, m)).astype(np.uint32)
            | (rng.integers(0, 2, size=(n, m)).astype(np.uint32) << 16))


def _online(seq, lags):
    """The accumulator as the reference describes it, value by value (floats, ring buffer)."""
    ring = [0.0] * max(lags)
    acc = {lag: [0, 0.0, 0.0, 0.0, 0.0, 0.0] for lag in lags}
    for n_obs, value in enumerate(seq):
        for lag in lags:
            if n_obs >= lag:
                earlier = ring[(n_obs - lag) % len(ring)]
                a = acc[lag]
                a[0] += 1
                a[1] += earlier
                a[2] += value
                a[3] += earlier * earlier
                a[4] += value * value
                a[5] += earlier * value
        ring[n_obs % len(ring)] = float(value)
    return acc


@pytest.mark.parametrize("lags", [(1,), (1, 2, 5), (3, 7)])
def test_state_matches_online_accumulator_and_joins(lags):
    rng = np.random.default_rng(5)
    obs = _random_obs(rng, 7, 40)
    whole = rd.StrategyLagState.from_observations(lags, obs)
    for i in range(obs.shape[0]):
        wins = _online([float(v >> 16) for v in obs[i]], lags)
        rounds = _online([float(v & 0xFFFF) for v in obs[i]], lags)
        for z, lag in enumerate(lags):
            assert whole.stats[i, z, 0] == wins[lag][0] == rounds[lag][0]
            assert whole.stats[i, z, 1:6].tolist() == wins[lag][1:]
            assert whole.stats[i, z, 6:11].tolist() == rounds[lag][1:]
    # any way of cutting the run into launches joins back to the same state, short pieces included
    for cuts in ([13], [1, 2, 3], [5, 6, 8, 9, 31], [39], list(range(1, 40))):
        bounds = [0, *cuts, obs.shape[1]]
        state = rd.StrategyLagState.empty(obs.shape[0], lags)
        for a, b in zip(bounds, bounds[1:]):
            state = state.extend(rd.StrategyLagState.from_observations(lags, obs[:, a:b]))
        assert state.n_obs == whole.n_obs
        assert np.array_equal(state.stats, whole.stats)
        assert np.array_equal(state.head, whole.head) and np.array_equal(state.tail, whole.tail)


def test_rows_shape_and_statuses():
    lags = (1, 3)
    obs = np.full((2, 6), 10, dtype=np.uint32)              # constant n_rounds, never a win
    obs[1] = [5 | 1 << 16, 7, 9 | 1 << 16, 4, 12, 6 | 1 << 16]
    state = rd.StrategyLagState.from_observations(lags, obs)
    rows = state.rows([40, 41], 3)
    assert len(rows) == 2 * 2 * 2
    assert [r["metric"] for r in rows[:4]] == ["win_indicator", "win_indicator", "n_rounds", "n_rounds"]
    assert {r["estimability_status"] for r in rows[:4]} == {"zero_variance"}
    assert all(r["autocorr"] is None for r in rows[:4])
    assert rows[4]["strategy"] == 41 and rows[4]["observations"] == 6 and rows[4]["lagged_pairs"] == 5
    assert rows[4]["estimability_status"] == "estimated" and -1.0 <= rows[4]["autocorr"] <= 1.0
    assert rows[5]["lagged_pairs"] == 3
    assert rows[4]["zero_centered_descriptive_reference_band_upper"] == 1.96 / 5**0.5
    assert rd.StrategyLagState.from_observations((4,), obs[:, :5]).rows([0, 1], 2) == []   # < lag + 2
    short = rd.StrategyLagState.from_observations((4,), obs)                                # 2 pairs
    assert short.rows([0, 1], 2)[0]["lagged_pairs"] == 2
    assert rd.normalize_lags(None) == (1,) and rd.normalize_lags([5, 1, 5, 0, -2]) == (1, 5)


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_rows_equal_reference_rows(monkeypatch):
    monkeypatch.syspath_prepend(str(REF))
    import farkle.analysis.rng_diagnostics as ref

    try:
        lags = (1, 2, 6)
        rng = np.random.default_rng(11)
        obs = _random_obs(rng, 5, 57)
        obs[3] = 20                                           # a zero-variance group
        mine = rd.StrategyLagState.from_observations(lags, obs).rows([9, 4, 7, 1, 3], 4)
        want = []
        for i, sid in enumerate([9, 4, 7, 1, 3]):
            rounds, wins = ref._OnlineMetric(lags), ref._OnlineMetric(lags)
            for v in obs[i]:
                rounds.push(float(v & 0xFFFF))
                wins.push(float(v >> 16))
            want += ref._rows_for_online_group((ref._GROUP_STRATEGY, 4, sid), lags=lags, rounds=rounds,
                                               wins=wins)
        assert mine == want                                   # floats compared by value: bit-identical
        assert rd.normalize_lags([3, 1, 3]) == ref._normalize_lags([3, 1, 3])
        assert rd.EXPECTED_NOTE == ref._EXPECTED_NOTE
        import pyarrow as pa

        assert rd.diagnostics_schema().equals(ref._stats_schema(), check_metadata=True)
        assert rd.diagnostics_table(mine).equals(pa.Table.from_pylist(want, schema=ref._stats_schema()))
    finally:
        for name in [m for m in sys.modules if m == "farkle" or m.startswith("farkle.")]:
            sys.modules.pop(name)


def _expected_state(z, lags):
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    obs = rd.observations_from_rows(z["rows"], len(z["strategies"]), nsh)
    return rd.StrategyLagState.from_observations(lags, obs)


def test_cell_in_batches_oracle_engine():
    """`strategy_lag_state` over launches of 3 shuffles == the whole fixture at once (CPU oracle)."""
    z = np.load(GOLDEN / "games_fast_54_4.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    lags = (1, 2, 4)
    want = _expected_state(z, lags)
    got = rd.strategy_lag_state(root, k, sh0, nsh, z["strategies"], lags, batch_shuffles=3,
                                engine=OracleEngine())
    assert got.n_obs == nsh and np.array_equal(got.stats, want.stats)
    assert np.array_equal(got.head, want.head) and np.array_equal(got.tail, want.tail)


@pytest.mark.gpu
@pytest.mark.parametrize("name,lags", [("fast_54_4", (1, 2, 4)), ("fast_42_2", (1,)), ("full_0_5", (2, 3))])
def test_device_lag_statistics(name, lags):
    from farkle_ii_b200.device import get_engine

    eng = get_engine(0)
    z = np.load(GOLDEN / f"games_{name}.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    want = _expected_state(z, lags)
    res = eng.play_tournament(root, k, sh0, nsh, z["strategies"], lags=lags)
    got = rd.StrategyLagState.from_launch(lags, nsh, res.lag_stats, res.lag_edges)
    assert res.lag_stats.shape == (len(z["strategies"]), len(lags), LAG_WIDTH)
    assert np.array_equal(got.stats, want.stats)
    assert np.array_equal(got.head, want.head) and np.array_equal(got.tail, want.tail)
    if nsh >= 3:                                              # launches of uneven size join up
        parts = rd.strategy_lag_state(root, k, sh0, nsh, z["strategies"], lags, batch_shuffles=max(nsh // 3, 1),
                                      engine=eng)
        assert np.array_equal(parts.stats, want.stats) and np.array_equal(parts.tail, want.tail)
    # the tallies of the same launch are untouched by the extra kernels
    plain = eng.play_tournament(root, k, sh0, nsh, z["strategies"])
    assert np.array_equal(res.tallies.cpu().numpy(), plain.tallies.cpu().numpy())


@pytest.mark.gpu
def test_device_lag_statistics_large_cell():
    """Full grid, 300 shuffles of k=2: device sums == brute force over the device's own rows
    (rows are bit-exact with the oracle elsewhere); two halves join to the whole."""
    from farkle_ii_b200.device import get_engine

    eng = get_engine(0)
    table = np.load(GOLDEN / "games_full_0_2.npz")["strategies"]
    lags, nsh = (1, 7, 64), 300
    res = eng.play_tournament(3, 2, 10, nsh, table, lags=lags, want_rows=True)
    obs = rd.observations_from_rows(res.rows_numpy(), len(table), nsh)
    want = rd.StrategyLagState.from_observations(lags, obs)
    got = rd.StrategyLagState.from_launch(lags, nsh, res.lag_stats, res.lag_edges)
    assert np.array_equal(got.stats, want.stats)
    assert np.array_equal(got.head, want.head) and np.array_equal(got.tail, want.tail)
    halves = rd.strategy_lag_state(3, 2, 10, nsh, table, lags, batch_shuffles=170, engine=eng)
    assert np.array_equal(halves.stats, want.stats)
    rows = got.rows(range(len(table)), 2)
    assert len(rows) == len(table) * 2 * len(lags)
    assert sum(r["estimability_status"] == "estimated" for r in rows) > len(rows) // 2


@pytest.mark.gpu
def test_lag_argument_errors():
    from farkle_ii_b200 import _native
    from farkle_ii_b200.device import get_engine

    eng = get_engine(0)
    table = np.load(GOLDEN / "games_fast_42_2.npz")["strategies"]
    for bad in ((0,), (1, 1), (5000,), tuple(range(1, 10))):
        with pytest.raises(_native.NativeError):
            eng.play_tournament(1, 2, 0, 4, table, lags=bad)
    # an empty launch leaves empty outputs, not stale memory
    res = eng.play_tournament(1, 2, 0, 0, table, lags=(1, 3), matchup_min_observations=3, want_first_seen=True)
    assert (res.first_seen.cpu().numpy() == -1).all() and len(res.matchup_count) == 0
    state = rd.StrategyLagState.from_launch((1, 3), 0, res.lag_stats, res.lag_edges)
    assert state.n_obs == 0 and not state.stats.any() and state.rows(range(len(table)), 2) == []
    one = rd.strategy_lag_state(1, 2, 0, 5, table, (1, 3), engine=eng)
    assert np.array_equal(state.extend(one).stats, one.stats)


# ---- matchup groups ------------------------------------------------------------------------------
def _rows_like(k, seats, rounds):
    """Minimal compact-row look-alike (seat strategies + n_rounds) for the host brute force."""
    rows = np.zeros(len(rounds), dtype=[("n_rounds", "<u2"), ("seats", [("strategy", "<i4")], (k,))])
    rows["n_rounds"] = rounds
    rows["seats"]["strategy"] = seats
    return rows


def test_matchup_groups_brute_force_and_selection():
    rng = np.random.default_rng(3)
    k, lags = 3, (1, 2)
    seats = np.array([rng.permutation(6)[:k] for _ in range(400)], dtype=np.int32)   # 20 possible matchups
    rounds = rng.integers(5, 60, size=400)
    groups = rd.MatchupLagGroups.from_rows(lags, _rows_like(k, seats, rounds), 3)
    assert len(groups) == 20 and int(groups.counts.sum()) == 400
    for g in range(len(groups)):
        mine = [int(r) for s, r in zip(seats, rounds) if sorted(s) == groups.participants[g].tolist()]
        acc = _online([float(v) for v in mine], lags)
        for z, lag in enumerate(lags):
            assert groups.stats[g, z].tolist() == acc[lag]
    few = rd.MatchupLagGroups.from_rows(lags, _rows_like(k, seats[:30], rounds[:30]), 3)
    assert 0 < len(few) < 20 and few.counts.min() >= 3
    # the cap keeps the groups with the smallest priorities, across cells
    other = rd.MatchupLagGroups.from_rows(lags, _rows_like(2, seats[:, :2], rounds), 3)
    masks = rd.select_matchup_groups([groups, other], 4, cap=9)
    assert sum(int(m.sum()) for m in masks) == 9
    pri = np.concatenate([groups.priorities(4), other.priorities(4)])
    kept = np.concatenate(masks)
    assert pri[kept].max() < pri[~kept].min()
    assert all(m.all() for m in rd.select_matchup_groups([groups, other], 4, cap=None))
    rows = groups.rows(4, masks[0])
    assert len(rows) == int(masks[0].sum()) * len(lags)
    assert rows[0]["summary_level"] == "matchup" and rows[0]["metric"] == "n_rounds"
    assert rows[0]["matchup"] == " | ".join(str(v) for v in rows[0]["participant_strategy_ids"])


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")
def test_matchup_identity_priority_and_rows_equal_reference(monkeypatch):
    monkeypatch.syspath_prepend(str(REF))
    import farkle.analysis.rng_diagnostics as ref

    try:
        rng = np.random.default_rng(8)
        k, width, lags = 3, 5, (1, 3)
        seats = np.array([rng.permutation(7)[:k] for _ in range(300)], dtype=np.int32)
        rounds = rng.integers(5, 60, size=300)
        groups = rd.MatchupLagGroups.from_rows(lags, _rows_like(k, seats, rounds), 4)
        padded = np.full((len(groups), width), -1, dtype=np.int32)
        padded[:, :k] = groups.participants
        want_ids = ref._matchup_ids(np.full(len(groups), k, dtype=np.int16), padded)
        assert np.array_equal(groups.group_ids(width), want_ids)
        records = np.zeros(len(groups), dtype=ref._count_dtype(width))
        records["group_type"], records["k"], records["group_id"] = ref._GROUP_MATCHUP, k, want_ids
        assert np.array_equal(groups.priorities(width), ref._priority(records))
        # selection order: the reference sorts (priority, group_type, k, group_id, p...) and cuts at the cap
        top = np.zeros(len(groups), dtype=ref._priority_key_dtype(width))
        top["priority"], top["group_type"], top["k"] = ref._priority(records), ref._GROUP_MATCHUP, k
        top["group_id"] = want_ids
        for c in range(width):
            top[f"p{c}"] = padded[:, c]
        cap = 11
        want_keep = np.zeros(len(groups), dtype=bool)
        want_keep[ref._priority_sort_order(top)[:cap]] = True
        (mask,) = rd.select_matchup_groups([groups], width, cap=cap)
        assert np.array_equal(mask, want_keep)
        # report rows
        mine = groups.rows(width, mask)
        want = []
        chosen = np.flatnonzero(mask)
        for g in chosen[np.lexsort((*(groups.participants[chosen, c] for c in reversed(range(k))),
                                    want_ids[chosen]))]:
            acc = ref._OnlineMetric(lags)
            for s, r in zip(seats, rounds):
                if sorted(s) == groups.participants[g].tolist():
                    acc.push(float(r))
            want += ref._rows_for_online_group((ref._GROUP_MATCHUP, k, int(want_ids[g]), *padded[g].tolist()),
                                               lags=lags, rounds=acc, wins=None)
        assert mine == want
    finally:
        for name in [m for m in sys.modules if m == "farkle" or m.startswith("farkle.")]:
            sys.modules.pop(name)


@pytest.mark.gpu
@pytest.mark.parametrize("k,n_strat,nsh,min_obs,lags", [
    (2, 80, 120, 3, (1, 2)),        # fast grid: 3,160 possible pairs, 4,800 games, exact keys
    (4, 8, 150, 3, (1, 5)),         # 70 possible matchups, long groups, hashed keys
    (3, 12, 40, 1, (2,)),           # every group eligible, many of size 1
    (6, 12, 60, 2, (1,)),
])
def test_device_matchup_groups(k, n_strat, nsh, min_obs, lags):
    from farkle_ii_b200.device import get_engine

    eng = get_engine(0)
    table = np.load(GOLDEN / "games_fast_42_2.npz")["strategies"][:n_strat]
    res = eng.play_tournament(21, k, 4, nsh, table, lags=lags, matchup_min_observations=min_obs,
                              want_rows=True, target_score=3000)
    got = rd.MatchupLagGroups.from_launch(lags, res)
    want = rd.MatchupLagGroups.from_rows(lags, res.rows_numpy(), min_obs)
    assert len(got) == len(want) > 0
    assert np.array_equal(got.participants, want.participants)
    assert np.array_equal(got.counts, want.counts)
    assert np.array_equal(got.stats, want.stats)
    # strategy ids other than the table position flow into the participants
    ids = (np.arange(n_strat, dtype=np.int32)[::-1] * 3 + 7).copy()
    res2 = eng.play_tournament(21, k, 4, nsh, table, strategy_ids=ids, lags=lags,
                               matchup_min_observations=min_obs, want_rows=True, target_score=3000,
                               strategy_lags=False)
    assert res2.lag_stats is None
    got2 = rd.MatchupLagGroups.from_launch(lags, res2)
    want2 = rd.MatchupLagGroups.from_rows(lags, res2.rows_numpy(), min_obs)
    assert np.array_equal(got2.participants, want2.participants) and np.array_equal(got2.stats, want2.stats)


@pytest.mark.gpu
def test_device_matchup_groups_full_grid_k2():
    """Full 5,160-strategy grid, k=2, 400 shuffles (1.03 M games over 13.3 M possible pairs):
    the groups with >= 3 games, checked against a numpy grouping of the device's rows."""
    from farkle_ii_b200.device import get_engine

    eng = get_engine(0)
    table = np.load(GOLDEN / "games_full_0_2.npz")["strategies"]
    lags, nsh = (1,), 400
    res = eng.play_tournament(5, 2, 0, nsh, table, lags=lags, matchup_min_observations=3, want_rows=True)
    got = rd.MatchupLagGroups.from_launch(lags, res)
    rows = res.rows_numpy()
    ids = np.sort(rows["seats"]["strategy"], axis=1).astype(np.int64)
    key = ids[:, 0] * len(table) + ids[:, 1]
    order = np.argsort(key, kind="stable")
    uniq, start, count = np.unique(key[order], return_index=True, return_counts=True)
    keep = count >= 3
    assert len(got) == int(keep.sum()) > 0
    assert np.array_equal(got.participants[:, 0] * len(table) + got.participants[:, 1], uniq[keep])
    assert np.array_equal(got.counts, count[keep])
    r = rows["n_rounds"].astype(np.int64)[order]
    for g in np.flatnonzero(keep)[:: max(int(keep.sum()) // 200, 1)]:      # a sample of the groups
        seq = r[start[g]:start[g] + count[g]]
        x, y = seq[:-1], seq[1:]
        want = [len(x), x.sum(), y.sum(), (x * x).sum(), (y * y).sum(), (x * y).sum()]
        at = int(np.searchsorted(uniq[keep], uniq[g]))
        assert got.stats[at, 0].tolist() == want


# ---- the whole stage against the reference's own output ---------------------------------------------
def _stage_rows(engine):
    """rng_diagnostics.parquet of the tiny run of tests/golden/make_golden_rngdiag.py, rebuilt from
    the tournament reduction: (rows, fixture)."""
    import json

    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    fixture = json.loads((GOLDEN / "rng_diagnostics_fast42.json").read_text())
    strategies = generate_strategy_grid(
        score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
        consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
        run_up_score_opts=[True])[0]
    assert len(strategies) == fixture["n_strategies"]
    n_shuffles = fixture["games"] // sum(len(strategies) // k for k in fixture["ks"])
    rows = rd.diagnose_root(fixture["root_seed"], fixture["ks"], n_shuffles, pack_strategies(strategies),
                            [s.strategy_id for s in strategies], fixture["lags"],
                            max_players=fixture["seat_strategy_columns"],
                            cap=fixture["summary"]["effective_matchup_group_cap"], engine=engine)
    return rows, fixture


def _canonical(rows):
    return sorted(rows, key=lambda r: (r["summary_level"], r["n_players"], r["strategy"] or 0,
                                       r["matchup_id"] or 0, r["metric"], r["lag"]))


def _check_stage(engine):
    rows, fixture = _stage_rows(engine)
    want = [dict(r, note=fixture["note"], sequence_order=fixture["sequence_order"][r["summary_level"]])
            for r in fixture["rows"]]
    assert len(rows) == len(want) == 648
    assert _canonical(rows) == _canonical(want)        # every value, floats bit for bit
    summary = fixture["summary"]
    assert sum(r["summary_level"] == "matchup" for r in rows) == \
        summary["eligible_matchup_group_count"] * len(fixture["lags"])
    assert sum(r["summary_level"] == "strategy" for r in rows) == \
        summary["eligible_strategy_group_count"] * 2 * len(fixture["lags"])


def test_whole_stage_equals_reference_output_oracle_engine():
    """Strategy AND matchup rows of the reference's `rng_diagnostics.parquet` (its real stage, run on
    rows its real runner wrote: tests/golden/make_golden_rngdiag.py) == the rows built from the lag
    sums of the reduction.  Compute served by the CPU oracle here."""
    _check_stage(OracleEngine())


@pytest.mark.gpu
def test_whole_stage_equals_reference_output_device():
    from farkle_ii_b200.device import get_engine

    _check_stage(get_engine(0))
