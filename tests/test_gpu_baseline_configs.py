"""GPU parity at the sizes BASELINE.json's configs name, the error paths, and a stress test of the
shared-memory staging ("What's weak" items 1, 9, 10).

All calls go through the C ABI; the checker is the oracle (pinned to the reference by
tests/test_oracle_golden.py).  Bar: bit-exact.
"""

from __future__ import annotations

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as fo  # noqa: E402  (test infrastructure: the checker)

THREADS = os.cpu_count() or 8
SHUFFLES, PER_BATCH = 4300, 43  # workload_planner.py:158-178 for the full grid


@pytest.fixture(scope="module")
def eng():
    from farkle_ii_b200.device import get_engine

    return get_engine(0)


@pytest.fixture(scope="module")
def full_grid(golden_dir):
    return np.load(golden_dir / "games_full_0_2.npz")["strategies"]


def _whole_cell_slots(eng, root, k, table, slots):
    """Play the WHOLE 4,300-shuffle cell in one launch with one tally slot per deterministic batch
    (the launch shape of the bench and of production) and check the given slots against the
    oracle's own run of those 43 shuffles; returns the slot tensor and the totals."""
    res = eng.play_tournament(root, k, 0, SHUFFLES, table, shuffles_per_slot=PER_BATCH)
    tallies, totals = res.tallies.cpu().numpy(), res.totals.cpu().numpy()
    assert tallies.shape[0] == SHUFFLES // PER_BATCH
    for b in slots:
        want_t, want_tot, _ = fo.play_tournament(root, k, b * PER_BATCH, PER_BATCH, table, n_threads=THREADS)
        assert np.array_equal(tallies[b], want_t[0]), (root, k, b)
        assert want_tot[0] == PER_BATCH * (len(table) // k)
    # size-independent properties of the whole cell (run_tournament.py:721-726)
    games = SHUFFLES * (len(table) // k)
    whole = tallies.sum(axis=0)
    assert totals[0] == games and totals[1] + totals[2] == games and totals[7] == 0
    assert (whole[:, 1] == SHUFFLES).all() and (whole[:, 2] + whole[:, 3] == SHUFFLES).all()
    assert whole[:, 0].sum() == totals[1] and whole[:, 3].sum() == k * totals[2]
    assert totals[8:8 + k].sum() == totals[1] and not totals[8 + k:].any()
    return tallies, totals


def test_config2_k6_full_size_cell(eng, full_grid):
    """BASELINE configs[2]: 6-player games over the full grid, root 42, all 4,300 shuffles
    (3,698,000 games, the "maximal lane divergence" config): three random deterministic batches
    bit-exact against the oracle, taken out of the one full-size launch."""
    rng = np.random.Generator(np.random.PCG64DXSM(606))
    slots = sorted(int(b) for b in rng.choice(SHUFFLES // PER_BATCH, size=3, replace=False))
    _whole_cell_slots(eng, 42, 6, full_grid, slots)


def test_config1_k4_full_size_cell(eng, full_grid):
    """BASELINE configs[1], the k=4 half of the bench step (root 43; the k=2 half is compared whole in
    test_gpu_parity.py::test_full_size_cell_bit_exact): five random batches of the full-size launch."""
    rng = np.random.Generator(np.random.PCG64DXSM(404))
    slots = sorted(int(b) for b in rng.choice(SHUFFLES // PER_BATCH, size=5, replace=False))
    _whole_cell_slots(eng, 43, 4, full_grid, slots)


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 8, 10, 12])
def test_config4_mega_root_every_k(eng, full_grid, k):
    """BASELINE configs[4] (configs/farkle_mega_config.yaml:10-14): root 102, every k of the mega
    config at its full 4,300 shuffles; one random deterministic batch per k against the oracle."""
    rng = np.random.Generator(np.random.PCG64DXSM(102_000 + k))
    _whole_cell_slots(eng, 102, k, full_grid, [int(rng.integers(0, SHUFFLES // PER_BATCH))])


def test_config3_h2h_production_blocks(eng, full_grid):
    """BASELINE configs[3]: H2H blocks at the production plan n_completed_required = 2,191,
    max_attempts = 4,382 (docs/remediation/task5a_production_capacity_report.md:58-61), chunk bound
    5,000 (h2h_schedule.py:94-129): 200 blocks over pairs of the first 150 grid ids, both orders,
    through `h2h.simulate_blocks` (launch, resolve, relaunch for replacements) against the oracle's
    attempt-by-attempt loop.  Pairs with a never-banking strategy exercise the early-stop boundary
    from both sides: blocks that stop exactly at `required` completed games, blocks that need
    replacement attempts, and blocks that run into `max_attempts`."""
    from farkle_ii_b200 import h2h
    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    strategies = generate_strategy_grid()[0]
    assert np.array_equal(pack_strategies(strategies), full_grid)     # ids = positions in the default grid
    manifest = h2h.build_strategy_manifest(strategies)
    required, max_attempts = 2191, 4382
    rng = np.random.Generator(np.random.PCG64DXSM(2191))
    never = [i for i in range(len(full_grid))
             if (full_grid["flags"][i] & 0x08) and full_grid["dice_threshold"][i] <= 0
             and (not (full_grid["flags"][i] & 0x04) or (full_grid["flags"][i] & 0x10))]
    assert never, "the default grid contains never-banking strategies (SURVEY.md section 7)"
    blocks = []
    for b in range(200):
        a, c = (int(x) for x in rng.choice(150, size=2, replace=False))
        if b % 25 == 0:      # safety-limit games: replacements and the max_attempts stop
            a = never[b % len(never)]
        if b % 50 == 0:
            c = never[(b + 1) % len(never)]
        blocks.append({"block_id": f"b{b}", "root_seed": 42, "pair_id": 1000 + b // 2, "order": b % 2,
                       "seat1_strategy": a, "seat2_strategy": c, "n_completed_required": required,
                       "max_attempts": max_attempts})
    def check(group, profile, max_rounds):
        got = h2h.simulate_blocks(group, manifest, 5000, profile)
        kinds = {}
        for blk, g in zip(group, got):
            prog, _ = fo.play_h2h_block(42, blk["pair_id"], blk["order"], full_grid[blk["seat1_strategy"]],
                                        full_grid[blk["seat2_strategy"]], n_completed_required=required,
                                        max_attempts=max_attempts, chunk_games=5000, max_rounds=max_rounds,
                                        progress=np.zeros(5, dtype=np.int32))
            mine = [g["games_attempted"], g["games_completed"], g["games_safety_limit"], g["wins_seat1"],
                    g["wins_seat2"]]
            assert mine == prog.tolist(), blk["block_id"]
            assert mine[0] == mine[1] + mine[2] and mine[3] + mine[4] == mine[1] <= required
            kind = "exact" if mine[0] == required else ("capped" if mine[1] < required else "replaced")
            kinds[kind] = kinds.get(kind, 0) + 1
        return kinds

    # production limits (200 rounds): a pair either always completes or (two never-banking seats) never
    kinds = check(blocks, None, 200)
    assert set(kinds) == {"exact", "capped"} and kinds["capped"] == 4, kinds
    # the same plan under a 22-round safety limit (mean game length ~24 rounds): most attempts of a
    # block hit the limit, so blocks need replacement attempts and stop inside the replacement range
    from farkle_ii_b200.limits import GameProfile

    kinds = check(blocks[1:25], GameProfile(default_max_rounds=22), 22)
    assert "replaced" in kinds, kinds


# ----------------------------------------------------------------------------------- error paths
def _never_banking_pair():
    t = np.zeros(2, dtype=fo.STRATEGY_DTYPE)
    t["score_threshold"], t["dice_threshold"] = 200, 0
    t["flags"] = 0x04 | 0x08 | 0x10            # consider score and dice, require both
    return t


def test_roll_limit_error_path(eng, full_grid, monkeypatch):
    """ROLL_LIMIT (engine.py:36,242-243 raises RuntimeError): unreachable in real play, so the test
    knob FB_TEST_ROLL_LIMIT lowers it for both the CUDA library and the oracle.  The same games
    carry FB_ROW_ROLL_LIMIT, `totals[7]` counts them, every other row is bit-identical, and the
    Python surface raises."""
    from farkle_ii_b200.layout import ROW_ROLL_LIMIT

    monkeypatch.setenv("FB_TEST_ROLL_LIMIT", "4")
    table = full_grid[:120]
    for k in (2, 3):
        res = eng.play_tournament(5, k, 0, 20, table, want_rows=True)
        want_t, want_tot, want_rows = fo.play_tournament(5, k, 0, 20, table, want_rows=True)
        rows, tot = res.rows_numpy(), res.totals.cpu().numpy()
        bad = (rows["flags"] & ROW_ROLL_LIMIT) != 0
        assert np.array_equal(bad, (want_rows["flags"] & ROW_ROLL_LIMIT) != 0)
        assert 0 < bad.sum() < len(rows) and tot[7] == want_tot[7] == bad.sum()
        as_bytes = lambda a: a.view(np.uint8).reshape(len(a), -1)  # noqa: E731  (row padding included)
        assert np.array_equal(as_bytes(rows)[~bad], as_bytes(want_rows)[~bad])
    monkeypatch.delenv("FB_TEST_ROLL_LIMIT")
    res = eng.play_tournament(5, 2, 0, 20, table, want_rows=True)      # and the knob is really off again
    assert res.totals.cpu().numpy()[7] == 0


def test_int16_overflow_flag_and_bounds(eng):
    """Counters that leave the int16 range of the row schema (utils/schema_helpers.py:44-60): two
    never-banking seats playing 30,000 rounds roll more than 32,767 times each.  The row carries
    FB_ROW_I16_OVERFLOW exactly as the oracle's does; max_rounds above FB_MAX_ROUNDS is rejected
    (the header keeps n_rounds in 16 bits)."""
    from farkle_ii_b200 import _native
    from farkle_ii_b200.layout import MAX_ROUNDS, ROW_I16_OVERFLOW, ROW_SAFETY_LIMIT

    table = _never_banking_pair()
    res = eng.play_tournament(1, 2, 0, 1, table, max_rounds=30_000, want_rows=True)
    _, want_tot, want_rows = fo.play_tournament(1, 2, 0, 1, table, max_rounds=30_000, want_rows=True)
    rows = res.rows_numpy()
    assert rows.tobytes() == want_rows.tobytes()
    assert rows["flags"][0] == ROW_I16_OVERFLOW | ROW_SAFETY_LIMIT and rows["n_rounds"][0] == 30_000
    assert res.totals.cpu().numpy()[7] == want_tot[7] == 1
    # the largest legal value still works and reports its rounds exactly
    res = eng.play_tournament(1, 2, 0, 1, table, max_rounds=MAX_ROUNDS, want_rows=True)
    assert res.rows_numpy()["n_rounds"][0] == MAX_ROUNDS
    for call in (lambda: eng.play_tournament(1, 2, 0, 1, table, max_rounds=MAX_ROUNDS + 1),
                 lambda: eng.play_tournament(1, 2, 0, 1, table, max_rounds=70_000),
                 lambda: eng.play_tournament(1, 2, 0, 1, table, overrides=[(0, 0, 40_000)]),
                 lambda: eng.play_games(np.array([[103, 1, 2, 0, 0, 0, 0]], dtype=np.uint64), 2, table,
                                        max_rounds_v=[65_536]),
                 lambda: eng.play_h2h(1, [0], [0], table[:1], table[1:], [0], [1], max_rounds=MAX_ROUNDS + 1)):
        with pytest.raises(_native.NativeError):
            call()


def test_tally_id_validation(eng, full_grid):
    """Ids that would address outside the tally buffers are rejected."""
    from farkle_ii_b200 import _native

    table = full_grid[:40]
    ok = eng.play_tournament(3, 2, 0, 2, table, strategy_ids=np.arange(40, dtype=np.int32) + 5, n_tally_ids=45)
    assert ok.tallies.shape[1] == 45
    for kw in (dict(n_tally_ids=39),                                                  # implicit ids need 40
               dict(n_tally_ids=0),
               dict(strategy_ids=np.arange(40, dtype=np.int32) + 5, n_tally_ids=44),  # id 44 is out
               dict(strategy_ids=np.arange(40, dtype=np.int32) - 1, n_tally_ids=40)):  # id -1
        with pytest.raises(_native.NativeError):
            eng.play_tournament(3, 2, 0, 2, table, **kw)
    with pytest.raises(_native.NativeError):
        eng.run_tournament_host(3, 2, 0, 2, table, strategy_ids=np.arange(40, dtype=np.int32) + 1,
                                n_tally_ids=40)


def test_error_rows_raise_on_the_python_surface(eng, full_grid, monkeypatch):
    """`totals[7]` -> RuntimeError in `run_cell` (farkle_ii_b200/run_tournament.py), as the reference's
    engine raises (engine.py:242-243)."""
    from farkle_ii_b200 import run_tournament as frt
    from farkle_ii_b200.strategies import generate_strategy_grid

    strategies = generate_strategy_grid(score_thresholds=[250, 300, 350, 400], smart_five_opts=[True],
                                        smart_one_opts=[True], consider_score_opts=[True],
                                        consider_dice_opts=[True], auto_hot_dice_opts=[True],
                                        run_up_score_opts=[True])[0]
    cfg = frt.TournamentConfig(n_players=2, num_shuffles=6, n_strategies=len(strategies))
    frt._init_worker(strategies, cfg)
    tasks = [frt.ShuffleTask(42, 2, s, 0, 0) for s in range(6)]
    frt._run_chunk(tasks)                      # fine at the real ROLL_LIMIT
    monkeypatch.setenv("FB_TEST_ROLL_LIMIT", "2")
    with pytest.raises(RuntimeError, match="ROLL_LIMIT"):
        frt._run_chunk(tasks)


# ------------------------------------------------------------------------ staging stress test
@pytest.mark.parametrize("warps", [1, 7, 32])
def test_staging_stress_many_tiny_launches(eng, warps, monkeypatch):
    """Insurance for the hand-rolled cp.async / shared-memory staging of play_kernel (compute-
    sanitizer is closed on this pool): many tiny launches back to back with the resident warps per
    SM forced to 1, 7 and 32, k in {2, 3, 5, 12}, max_rounds in {1, 2, 200}, random tables, every
    game's row and the tallies against the oracle.  Short games maximise the number of game starts,
    turn switches, mispredicted next seats and queue refills per lane."""
    monkeypatch.setenv("FB_PLAY_WARPS", str(warps))
    rng = np.random.Generator(np.random.PCG64DXSM(7_000 + warps))
    cases = 200 if warps == 7 else 60
    for case in range(cases):
        k = int(rng.choice([2, 3, 5, 12]))
        n = k * int(rng.integers(1, 30))
        table = np.zeros(n, dtype=fo.STRATEGY_DTYPE)
        table["score_threshold"] = rng.integers(0, 30, size=n) * 50
        table["dice_threshold"] = rng.integers(-1, 7, size=n)
        flags = rng.integers(0, 256, size=n)
        flags &= np.where(flags & 0x01, 0xFF, 0xFF & ~0x02)
        flags &= np.where((flags & 0x0C) == 0x0C, 0xFF, 0xFF & ~0x10)
        table["flags"] = flags
        max_rounds = int(rng.choice([1, 2, 200]))
        target = int(rng.choice([300, 1500, 10_000]))
        nsh = int(rng.integers(1, 25))
        root, sh0 = int(rng.integers(0, 2**50)), int(rng.integers(0, 2**30))
        res = eng.play_tournament(root, k, sh0, nsh, table, target_score=target, max_rounds=max_rounds,
                                  want_rows=True)
        want_t, want_tot, want_rows = fo.play_tournament(root, k, sh0, nsh, table, target_score=target,
                                                         max_rounds=max_rounds, want_rows=True, n_threads=4)
        assert res.rows_numpy().tobytes() == want_rows.tobytes(), (warps, case, k)
        assert np.array_equal(res.tallies.cpu().numpy(), want_t), (warps, case, k)
        assert np.array_equal(res.totals.cpu().numpy(), want_tot), (warps, case, k)


# ------------------------------------------------------------------------ pipelined cell lists
def test_play_cells_equals_one_launch_per_cell(eng, full_grid):
    """`fb_play_tournament_cells` (cells pipelined over two workspace slots, preparation on a second
    stream, optional look-ahead cell) gives exactly what one `fb_play_tournament` per cell gives:
    mixed k, slotted tallies, accumulation into shared tensors, a look-ahead that is used, one that
    is not, and one that is superseded by a different first cell."""
    import torch

    from farkle_ii_b200.layout import TALLY_WIDTH, TOTALS_WIDTH

    n = len(full_grid)
    table = eng.to_device(full_grid)
    cells = [(7, 2, 0, 60), (7, 4, 10, 45), (8, 12, 3, 20), (7, 3, 500, 33), (9, 6, 0, 41), (9, 2, 60, 60)]

    def fresh():
        return (torch.zeros((2, n, TALLY_WIDTH), dtype=torch.int64, device=eng.device),
                torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=eng.device))

    want = []
    for root, k, s0, cnt in cells:
        t, tot = fresh()
        eng.play_tournament(root, k, s0, cnt, table, shuffles_per_slot=43, tallies=t, totals=tot)
        want.append((t.cpu(), tot.cpu()))
    # (a) the whole list in one call
    got = [fresh() for _ in cells]
    eng.play_cells([c + g for c, g in zip(cells, got)], table, shuffles_per_slot=43)
    for (t, tot), (wt, wtot) in zip(got, want):
        assert torch.equal(t.cpu(), wt) and torch.equal(tot.cpu(), wtot)
    # (b) cell by cell with the next one as look-ahead; the last look-ahead is never played, and the
    # call after it starts with a different cell
    got = [fresh() for _ in cells]
    for i, c in enumerate(cells):
        nxt = cells[i + 1] if i + 1 < len(cells) else (1234, 5, 0, 10)
        eng.play_cells([c + got[i]], table, ahead=nxt, shuffles_per_slot=43)
    again = fresh()
    eng.play_cells([cells[0] + again], table, shuffles_per_slot=43)
    for (t, tot), (wt, wtot) in zip(got + [again], want + [want[0]]):
        assert torch.equal(t.cpu(), wt) and torch.equal(tot.cpu(), wtot)
    # (c) two cells accumulating into the same tensors == the union; tallies-only and totals-only cells
    t, tot = fresh()
    eng.play_cells([(7, 2, 0, 30, t, tot), (7, 2, 30, 30, t, None), (7, 2, 30, 30, None, tot)], table,
                   shuffles_per_slot=0)
    assert torch.equal(t.cpu()[0], want[0][0].sum(dim=0)) and torch.equal(tot.cpu(), want[0][1])
    # against the oracle as well, through explicit ids
    ids = np.arange(n, dtype=np.int32)[::-1].copy()
    t = torch.zeros((1, n, TALLY_WIDTH), dtype=torch.int64, device=eng.device)
    eng.play_cells([(11, 5, 2, 9, t, None)], table, strategy_ids=ids)
    want_t, _, _ = fo.play_tournament(11, 5, 2, 9, full_grid, strategy_ids=ids, n_threads=THREADS)
    assert np.array_equal(t.cpu().numpy(), want_t)


def test_timeline_of_a_pipelined_cell_list(eng, full_grid):
    """`fb_timeline`: one mark per kernel of the tournament path, in stream order, the preparation of
    the next cell on its own lane and released by the end of the play kernel before it; nothing is
    recorded when the hook is off."""
    import torch

    from farkle_ii_b200.layout import TALLY_WIDTH, TOTALS_WIDTH

    n = len(full_grid)
    table = eng.to_device(full_grid)
    t = torch.zeros((1, n, TALLY_WIDTH), dtype=torch.int64, device=eng.device)
    tot = torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=eng.device)
    cells = [(3, 2, 0, 43, t, tot), (3, 4, 0, 43, t, tot), (3, 6, 0, 43, t, tot)]
    eng.play_cells(cells, table)  # streams and events exist afterwards
    torch.cuda.synchronize()
    eng.timeline(False)
    eng.play_cells(cells, table)
    assert eng.timeline_marks() == []
    eng.timeline(True)
    eng.play_cells(cells, table)
    marks = eng.timeline_marks()
    eng.timeline(False)
    main = [name for lane, name, _ in marks if lane == 0]
    prep = [name for lane, name, _ in marks if lane == 1]
    assert main == ["play_begin", "play_kernel", "finish", "gather"] * 3
    assert prep == ["prepare_begin", "permute", "seed"] * 3          # (the first cell is prepared on lane 1 too)
    at = {}
    for lane, name, ms in marks:
        at.setdefault((lane, name), []).append(ms)
    for lane in (0, 1):  # marks of one stream are in stream order
        times = [ms for ln, _name, ms in marks if ln == lane]
        assert times == sorted(times)
    # cell i+1 is prepared after cell i's play kernel has finished and before its own play kernel starts
    for i in (1, 2):
        assert at[(1, "prepare_begin")][i] >= at[(0, "play_kernel")][i - 1]
        assert at[(1, "seed")][i] <= at[(0, "play_kernel")][i]


# ------------------------------------------------------------------------ all-player statistics (f-3)
@pytest.mark.parametrize("name,spb,with_ids", [("fast_54_4", 2, False), ("fast_42_2", 5, True), ("full_0_5", 1, False),
                                               ("full_42_6", 2, False), ("full_102_12", 3, True)])
def test_all_player_statistics(eng, golden_dir, name, spb, with_ids):
    """`allplayer_gather_kernel` (optional output of the tournament launch): the unconditional
    all-player sufficient statistics per (deterministic batch, strategy) -- integer counts and sums,
    and the float64 sums of score / n_turns and score / n_rounds added in shuffle order -- equal, bit
    for bit, the host restatement over the same launch's rows.  That restatement and the Arrow table
    built from it are compared with the reference's own metrics stage in
    tests/test_reference_dropin.py::test_all_player_statistics_match_reference_metrics_stage."""
    from all_player_rows import all_player_from_rows

    z = np.load(golden_dir / f"games_{name}.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    table = z["strategies"]
    n = len(table)
    ids = (np.arange(n, dtype=np.int32)[::-1] * 2 + 1).copy() if with_ids else None
    res = eng.play_tournament(root, k, sh0, nsh, table, shuffles_per_slot=spb, want_all_player=True,
                              want_rows=True, want_game_seeds=True, strategy_ids=ids)
    rows = res.rows_numpy()
    gps = n // k
    batch = (np.arange(len(rows)) // gps) // spb
    n_ids = n if ids is None else int(ids.max()) + 1
    want = all_player_from_rows(rows, batch, -(-nsh // spb), n_ids)
    got = res.all_player.cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    # the float columns really are doubles with sensible values
    f = got[..., 41].view(np.float64)
    seated = got[..., 0] > 0
    assert np.isfinite(f).all() and (f[seated] >= 0).all() and seated.sum() == -(-nsh // spb) * n


# ------------------------------------------------------------------------ several devices, one process
def test_two_devices_in_one_process(full_grid):
    """Per-device library contexts ("re-entrant per (device, stream)"): two
    engines on two GPUs of one process, launches enqueued on both before either is waited for, give
    the two-rank result.  Needs two visible GPUs (skipped on the one-GPU box)."""
    import torch

    from farkle_ii_b200.device import get_engine

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    e0, e1 = get_engine(0), get_engine(1)
    assert e0 is not e1 and e0.device != e1.device
    table = full_grid[:240]
    a = e0.play_tournament(31, 4, 0, 40, table)
    b = e1.play_tournament(31, 4, 40, 40, table)
    c = e0.play_tournament(32, 2, 0, 10, table, want_rows=True)          # and back on the first device
    want_t, want_tot, _ = fo.play_tournament(31, 4, 0, 80, table, n_threads=THREADS)
    assert np.array_equal(a.tallies.cpu().numpy() + b.tallies.cpu().numpy(), want_t)
    assert np.array_equal(a.totals.cpu().numpy() + b.totals.cpu().numpy(), want_tot)
    _, _, want_rows = fo.play_tournament(32, 2, 0, 10, table, want_rows=True)
    assert c.rows_numpy().tobytes() == want_rows.tobytes()
    out = np.zeros((1, len(table), 26), dtype=np.int64)
    e1.run_tournament_host(31, 4, 0, 80, table, out_tallies=out)        # the host-buffer call on device 1
    assert np.array_equal(out, want_t)
