"""GPU parity: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bar: bit-exact.  Everything on this path is integer work.
"""

from __future__ import annotations

import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as fo  # noqa: E402  (test infrastructure: the checker)
from test_oracle_golden import GAME_FIXTURES, assert_rows_equal, check_expected_rows  # noqa: E402

P_PLAYER, P_SHUFFLE, P_PERM, P_GAME, P_TPLAYER, P_H2H_GAME, P_H2H_PLAYER = (
    10, 100, 101, 102, 103, 202, 203)


@pytest.fixture(scope="module")
def eng():
    from farkle_ii_b200.device import get_engine

    return get_engine(0)


def _coords9(c):
    return [c.get("purpose"), c["root_seed"], c.get("k", 0), c.get("shuffle_index", 0),
            c.get("pair_id", 0), c.get("order", 0), c.get("game_index", 0),
            c.get("seat_index", 0), c.get("replicate_index", 0)]


# --------------------------------------------------------------------------- RNG
def test_seedseq_words(eng, golden_dir):
    data = json.loads((golden_dir / "rng.json").read_text())
    for case in data["seedseq"]:
        got = eng.seedseq_generate(np.array(case["entropy"], dtype=np.uint32), 8)[0]
        assert got.tolist() == case["words"]
    # reference KAT tests/unit/utils/test_random_utils.py:32-40,73-78
    e = np.array([[1, 102, 32, 0, 2, 0, 194, 0, 0, 0, 0, 0, 18, 0, 0, 0, 0, 0],
                  [1, 102, 32, 0, 2, 0, 4052, 0, 0, 0, 0, 0, 4, 0, 0, 0, 0, 0]], dtype=np.uint32)
    assert eng.seedseq_generate(e, 1)[:, 0].tolist() == [2_963_478_802, 2_963_478_802]


def test_coordinate_streams_and_dice(eng, golden_dir):
    data = json.loads((golden_dir / "rng.json").read_text())
    pattern = data["n_dice_pattern"]
    coords = np.array([_coords9(c["coord"]) for c in data["coords"]], dtype=np.uint64)
    si = eng.seed_streams(coords)
    faces = eng.roll_dice(si, pattern)
    for i, case in enumerate(data["coords"]):
        assert (int(si[i, 0]) << 64) | int(si[i, 1]) == int(case["state"])
        assert (int(si[i, 2]) << 64) | int(si[i, 3]) == int(case["inc"])
        for r, n in enumerate(pattern):
            assert faces[i, r, :n].tolist() == case["dice"][r]
            assert not faces[i, r, n:].any()
        c = case["coord"]
        for as_u32, key in ((True, "seed_u32"), (False, "seed_u64")):
            got = eng.coordinate_seeds(c["purpose"], root_seed=c["root_seed"], k=c.get("k", 0),
                                       shuffle_index=c.get("shuffle_index", 0),
                                       pair_id=c.get("pair_id", 0), order=c.get("order", 0),
                                       game_index=0, vary="game_index",
                                       base=c.get("game_index", 0), n=1, as_u32=as_u32)
            if c.get("seat_index", 0) == 0:  # fingerprints never carry a seat
                assert int(got[0]) == case[key]


def test_coordinate_seed_ranges_vs_oracle(eng):
    got = eng.coordinate_seeds(P_SHUFFLE, root_seed=42, k=2, vary="shuffle_index", base=0, n=600,
                               as_u32=True)
    assert int(got[0]) == 1_998_876_487 and int(got[1]) == 2_468_994_662
    for i in (0, 17, 599):
        assert int(got[i]) == fo.coordinate_seed(P_SHUFFLE, root_seed=42, k=2, shuffle_index=i,
                                                 as_u32=True)
    got = eng.coordinate_seeds(P_H2H_GAME, root_seed=7, k=2, pair_id=5, order=1, vary="game_index",
                               base=2**33, n=50)
    for i in (0, 49):
        assert int(got[i]) == fo.coordinate_seed(P_H2H_GAME, root_seed=7, k=2, pair_id=5, order=1,
                                                 game_index=2**33 + i)


def test_lemire_rejection_branch(eng, golden_dir):
    data = json.loads((golden_dir / "rng.json").read_text())
    m = 2**64 - 1
    for case in data["rejection"]:
        state, inc = int(case["state"]), int(case["inc"])
        si = np.array([[state >> 64, state & m, inc >> 64, inc & m]], dtype=np.uint64)
        hb = np.array([[case.get("has32", 0), case.get("saved", 0)]], dtype=np.uint32)
        faces = eng.roll_dice(si, case["n_dice"], half_buffer=hb)[0]
        for r, n in enumerate(case["n_dice"]):
            assert faces[r, :n].tolist() == case["dice"][r]


def test_games_with_rejected_halves(eng, golden_dir):
    """The face queue of play_kernel keeps a half that NumPy's Lemire test rejects as a skip code
    (csrc/play.cuh).  tests/golden/rejects.json lists 18 tournament games of the full grid whose dice
    meet such a half (found with the oracle, four 32-bit values in 2^32): each is played alone at its
    explicit coordinates, and its whole shuffle through the tournament launch, rows / tallies / totals
    (incl. the count of 64-bit outputs the reference's generators would have produced) against the
    oracle."""
    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    table = pack_strategies(generate_strategy_grid()[0])
    games = json.loads((golden_dir / "rejects.json").read_text())["games"]
    assert len(games) >= 10
    by_k: dict[int, list] = {}
    for g in games:
        by_k.setdefault(g["k"], []).append(g)
    for k, gs in by_k.items():
        coords = np.array([[P_TPLAYER, g["root"], k, g["shuffle"], 0, 0, g["game"]] for g in gs], dtype=np.uint64)
        ids = np.stack([fo.permutation(g["root"], k, g["shuffle"], len(table))[g["game"] * k:(g["game"] + 1) * k]
                        for g in gs]).astype(np.int32)
        rows, totals = eng.play_games(coords, k, table[ids], seat_strategy_ids=ids)
        want_rows, want_tot = fo.play_games(coords, k, table[ids], seat_strategy_ids=ids)
        assert rows.tobytes() == want_rows.tobytes(), k
        assert np.array_equal(totals, want_tot), (k, totals[:8], want_tot[:8])
    for g in games[::3]:  # the whole shuffle of every third game through the tournament launch
        root, k, sh = g["root"], g["k"], g["shuffle"]
        tallies, totals, rows = _play(eng, root, k, sh, 1, table)
        want_t, want_tot, want_rows = fo.play_tournament(root, k, sh, 1, table, want_rows=True,
                                                         want_game_seeds=True, n_threads=1)
        assert rows.tobytes() == want_rows.tobytes(), g
        assert np.array_equal(tallies, want_t) and np.array_equal(totals, want_tot), g


def test_frequent_rejected_halves_variant():
    """The rejection paths of the face queue under load: a TEST BUILD of the same sources with the
    Lemire threshold raised to 2^30 (one half in four rejected; `_native.TEST_VARIANTS`), loaded by a
    subprocess through FARKLE_B200_LIB, against the oracle with the matching knob
    (tests/rejects_variant_check.py): rows, tallies, totals of k = 2, 3, 4, 5, 12 cells."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    from farkle_ii_b200 import _native

    lib = _native.build_variant("rejects")  # built by __graft_entry__.build(); kept when up to date
    env = dict(os.environ, FARKLE_B200_LIB=str(lib), FB_TEST_LEMIRE_THR="0x40000000")
    proc = subprocess.run([sys.executable, str(Path(__file__).parent / "rejects_variant_check.py")], env=env,
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert proc.stdout.count(": ok;") == 6, proc.stdout


def test_permutations(eng, golden_dir):
    perms = np.load(golden_dir / "perm.npz")
    for key in perms.files:
        root, k, sh, n = (int(x) for x in key.split("_"))
        assert np.array_equal(eng.permute_shuffles(root, k, sh, 1, n)[0], perms[key]), key
    got = eng.permute_shuffles(42, 5, 100, 37, 5160)
    for j in (0, 13, 36):
        assert np.array_equal(got[j], fo.permutation(42, 5, 100 + j, 5160))
    assert all(np.array_equal(np.sort(row), np.arange(5160)) for row in got)


# ----------------------------------------------------------------------- scoring
def _strategies(rows):
    s = np.zeros(len(rows), dtype=fo.STRATEGY_DTYPE)
    s["score_threshold"], s["dice_threshold"], s["flags"] = rows[:, 0], rows[:, 1], rows[:, 2]
    return s


def test_score_table_and_golden_rolls(eng, golden_dir):
    z = np.load(golden_dir / "scoring.npz")
    tab = z["table"]
    faces = np.zeros((len(tab), 6), dtype=np.uint8)
    for i, row in enumerate(tab):
        f = [face + 1 for face in range(6) for _ in range(int(row[face]))]
        faces[i, : len(f)] = f
    plain = np.zeros(len(tab), dtype=fo.STRATEGY_DTYPE)
    out = eng.default_score(faces, np.zeros(len(tab), dtype=np.int32), plain)
    n = (faces > 0).sum(axis=1)
    assert np.array_equal(out[:, 0], tab[:, 6]) and np.array_equal(out[:, 1], tab[:, 7])
    assert np.array_equal(out[:, 2], n - tab[:, 7]) and not out[:, 3:].any()
    rolls = z["csv_rolls"]  # tests/data/test_farkle_scores_data.csv of the reference
    out = eng.default_score(rolls[:, :6].astype(np.uint8), np.zeros(len(rolls), dtype=np.int32),
                            np.zeros(len(rolls), dtype=fo.STRATEGY_DTYPE))
    assert np.array_equal(out[:, :3], rolls[:, 6:9])


def test_default_score_sweep(eng, golden_dir):
    z = np.load(golden_dir / "scoring.npz")
    sin, want = z["sweep_in"], z["sweep_out"]
    out = eng.default_score(sin[:, :6].astype(np.uint8), sin[:, 6], _strategies(sin[:, 7:10]))
    assert np.array_equal(out, want)
    disc = z["discards"]
    faces = np.zeros((len(disc), 6), dtype=np.uint8)
    for i, row in enumerate(disc):
        f = [face + 1 for face in range(6) for _ in range(int(row[face]))]
        faces[i, : len(f)] = f
    flags = (disc[:, 12] | disc[:, 13] << 1 | disc[:, 9] << 2 | disc[:, 10] << 3 | disc[:, 11] << 4
             | 0x80)
    st = _strategies(np.stack([disc[:, 7], disc[:, 8], flags], axis=1))
    out = eng.default_score(faces, disc[:, 6], st)
    assert np.array_equal(out[:, 3:], disc[:, 14:16])


def test_default_score_exhaustive_vs_oracle(eng, golden_dir):
    """All 923 histograms x a parameter lattice, against the oracle's literal candidate search."""
    tab = np.load(golden_dir / "scoring.npz")["table"]
    rng = np.random.Generator(np.random.PCG64DXSM(7))
    faces, ts, st = [], [], []
    for row in tab:
        f = [face + 1 for face in range(6) for _ in range(int(row[face]))]
        for _ in range(40):
            flags = int(rng.integers(0, 256))
            if flags & 2 and not flags & 1:
                flags &= ~2
            if flags & 16 and (flags & 12) != 12:
                flags &= ~16
            faces.append(f + [0] * (6 - len(f)))
            ts.append(int(rng.integers(0, 30)) * 50)
            st.append((int(rng.choice([199, 200, 250, 300, 400, 550, 1000])),
                       int(rng.integers(-1, 6)), flags))
    faces = np.array(faces, dtype=np.uint8)
    st_arr = _strategies(np.array(st))
    out = eng.default_score(faces, np.array(ts, dtype=np.int32), st_arr)
    for i in range(0, len(faces), 7):  # the oracle is called per case; sample every 7th
        f = [int(x) for x in faces[i] if x]
        assert tuple(int(x) for x in out[i]) == fo.default_score(f, ts[i], st_arr[i:i + 1]), i


# ------------------------------------------------------------------- whole games
def _play(eng, root, k, sh0, nsh, table, **kw):
    res = eng.play_tournament(root, k, sh0, nsh, table, want_rows=True, want_game_seeds=True, **kw)
    return res.tallies.cpu().numpy(), res.totals.cpu().numpy(), res.rows_numpy()


@pytest.mark.parametrize("name", GAME_FIXTURES)
def test_tournament_rows_and_tallies_golden(eng, golden_dir, name):
    z = np.load(golden_dir / f"games_{name}.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    tallies, totals, rows = _play(eng, root, k, sh0, nsh, z["strategies"])
    assert_rows_equal(rows, z["rows"], k)
    assert np.array_equal(rows["game_ordinal"], np.arange(len(rows)))
    assert np.array_equal(tallies[0], z["tallies"])
    want_t, want_tot, want_rows = fo.play_tournament(root, k, sh0, nsh, z["strategies"],
                                                     want_rows=True, want_game_seeds=True)
    assert rows.tobytes() == want_rows.tobytes()
    assert np.array_equal(totals, want_tot)


def test_reference_raw_oracle_12_games(eng, golden_dir):
    z = np.load(golden_dir / "oracle12.npz")

    def play(root, k, overrides):
        tallies, _totals, rows = _play(eng, root, k, 0, 2, z["strategies"], target_score=100,
                                       overrides=overrides)
        assert_rows_equal(rows, z[f"rows_{root}_{k}"], k)
        assert np.array_equal(tallies[0], z[f"tallies_{root}_{k}"])
        return rows

    check_expected_rows(play)


def test_public_helper_kat(eng, golden_dir):
    # reference tests/unit/simulation/test_simulation.py:184-198
    z = np.load(golden_dir / "helpers.npz")
    coords = np.array([[P_PLAYER, 123, 3, 0, 0, 0, g] for g in range(10)], dtype=np.uint64)
    rows, totals = eng.play_games(coords, 3, np.tile(z["strategies"], (10, 1)), target_score=5000)
    assert np.bincount(rows["winner_seat"], minlength=3).tolist() == [2, 6, 2]
    want = z["rows"].copy()
    want["game_seed"] = 0
    assert_rows_equal(rows, want, 3)
    assert totals[0] == 10 and totals[8:11].tolist() == [2, 6, 2]


def test_fast_grid_seed42_full_cell(eng, golden_dir):
    z = np.load(golden_dir / "fast42.npz")
    res = eng.play_tournament(42, 2, 0, 600, z["strategies"])
    tallies, totals = res.tallies.cpu().numpy(), res.totals.cpu().numpy()
    assert np.array_equal(tallies[0], z["tallies"])
    assert totals[:3].tolist() == [24000, 23801, 199]
    assert tallies[0][[42, 46, 51, 37, 25], 0].tolist() == [330, 322, 407, 386, 140]


@pytest.mark.parametrize("k,nsh,spb", [(2, 43, 0), (3, 9, 4), (4, 12, 5), (5, 10, 0), (6, 7, 3),
                                       (8, 5, 0), (10, 4, 2), (12, 6, 0)])
def test_full_grid_cells_vs_oracle(eng, golden_dir, k, nsh, spb):
    """Full 5,160-strategy grid: tallies (per deterministic batch slot), totals and rows."""
    table = np.load(golden_dir / "games_full_0_2.npz")["strategies"]
    root, sh0 = 1000 + k, 4300 - nsh
    res = eng.play_tournament(root, k, sh0, nsh, table, shuffles_per_slot=spb, want_rows=True)
    want_t, want_tot, want_rows = fo.play_tournament(root, k, sh0, nsh, table,
                                                     shuffles_per_slot=spb, want_rows=True,
                                                     n_threads=8)
    tallies = res.tallies.cpu().numpy()
    assert np.array_equal(tallies, want_t)
    assert np.array_equal(res.totals.cpu().numpy(), want_tot)
    assert res.rows_numpy().tobytes() == want_rows.tobytes()
    # size-independent invariants (run_tournament.py:721-726): every strategy sits once per shuffle
    tot = tallies.sum(axis=0)
    assert (tot[:, 1] == nsh).all() and (tot[:, 2] + tot[:, 3] == tot[:, 1]).all()
    assert tot[:, 0].sum() == want_tot[1] and tot[:, 3].sum() == k * want_tot[2]


def test_strategy_id_mapping_and_accumulation(eng, golden_dir):
    table = np.load(golden_dir / "games_fast_42_2.npz")["strategies"]
    ids = (np.arange(80, dtype=np.int32) * 3 + 5)
    res = eng.play_tournament(9, 4, 0, 6, table, strategy_ids=ids, want_rows=True)
    want_t, _, want_rows = fo.play_tournament(9, 4, 0, 6, table, strategy_ids=ids, want_rows=True)
    assert np.array_equal(res.tallies.cpu().numpy(), want_t)
    assert res.rows_numpy().tobytes() == want_rows.tobytes()
    # two launches accumulated into the same tensors == one launch over the union
    a = eng.play_tournament(9, 4, 0, 3, table)
    eng.play_tournament(9, 4, 3, 3, table, tallies=a.tallies, totals=a.totals)
    whole, tot, _ = fo.play_tournament(9, 4, 0, 6, table)
    assert np.array_equal(a.tallies.cpu().numpy(), whole)
    assert np.array_equal(a.totals.cpu().numpy(), tot)


def test_edge_cases(eng, golden_dir):
    table = np.load(golden_dir / "games_fast_42_2.npz")["strategies"]
    # zero shuffles: nothing launched, zero tallies
    res = eng.play_tournament(1, 2, 0, 0, table)
    assert not res.tallies.cpu().numpy().any() and res.n_games == 0
    # k that does not divide the grid is rejected like _init_worker (run_tournament.py:274-275)
    from farkle_ii_b200._native import NativeError

    with pytest.raises(NativeError, match="n_players must divide"):
        eng.play_tournament(1, 3, 0, 1, table)
    with pytest.raises(NativeError):
        eng.play_tournament(1, 13, 0, 1, table)
    # single-seat games, max_rounds=1 (everything safety-limited unless 10k in one turn)
    for k, mr, tgt in ((1, 200, 10_000), (2, 1, 10_000), (5, 3, 2_000), (2, 200, 50)):
        res = eng.play_tournament(3, k, 5, 2, table, max_rounds=mr, target_score=tgt, want_rows=True)
        want_t, want_tot, want_rows = fo.play_tournament(3, k, 5, 2, table, max_rounds=mr,
                                                         target_score=tgt, want_rows=True)
        assert np.array_equal(res.tallies.cpu().numpy(), want_t), (k, mr, tgt)
        assert np.array_equal(res.totals.cpu().numpy(), want_tot)
        assert res.rows_numpy().tobytes() == want_rows.tobytes()


def test_play_games_explicit_vs_oracle(eng):
    rng = np.random.Generator(np.random.PCG64DXSM(99))
    n, k = 700, 3
    coords = np.zeros((n, 7), dtype=np.uint64)
    coords[:, 0] = rng.choice([P_PLAYER, P_TPLAYER, P_H2H_PLAYER], size=n)
    coords[:, 1] = rng.integers(0, 2**63, size=n)
    coords[:, 2] = k
    coords[:, 3] = rng.integers(0, 5000, size=n)
    coords[:, 4] = rng.integers(0, 2**40, size=n)
    coords[:, 5] = rng.integers(0, 2, size=n)
    coords[:, 6] = rng.integers(0, 2**34, size=n)
    st = np.zeros((n, k), dtype=fo.STRATEGY_DTYPE)
    st["score_threshold"] = rng.integers(1, 20, size=(n, k)) * 50
    st["dice_threshold"] = rng.integers(0, 5, size=(n, k))
    flags = rng.integers(0, 256, size=(n, k))
    flags &= np.where(flags & 1, 0xFF, ~2 & 0xFF)
    flags &= np.where((flags & 12) == 12, 0xFF, ~16 & 0xFF)
    st["flags"] = flags
    ids = rng.integers(0, 1000, size=(n, k)).astype(np.int32)
    tgt = rng.choice([100, 1000, 5000, 10_000], size=n).astype(np.int32)
    mr = rng.choice([0, 1, 5, 200], size=n).astype(np.int32)
    rows, totals = eng.play_games(coords, k, st, seat_strategy_ids=ids, target_scores=tgt,
                                  max_rounds_v=mr)
    want_rows, want_tot = fo.play_games(coords, k, st, seat_strategy_ids=ids, target_scores=tgt,
                                        max_rounds_v=mr)
    assert rows.tobytes() == want_rows.tobytes()
    assert np.array_equal(totals, want_tot)


# -------------------------------------------------------------------------- H2H
def test_h2h_blocks_golden(eng, golden_dir):
    blocks = json.loads((golden_dir / "h2h.json").read_text())
    keys = ("games_attempted", "games_completed", "games_safety_limit", "wins_seat1", "wins_seat2")

    def st(b, name):
        s = np.zeros(len(b), dtype=fo.STRATEGY_DTYPE)
        for i, blk in enumerate(b):
            s[i] = tuple(blk[name])
        return s

    pair = [b["pair_id"] for b in blocks]
    order = [b["order"] for b in blocks]
    req = [b["n_completed_required"] for b in blocks]
    # chunk of 13 attempts, as in the fixture (stop = min(max_attempts, attempted + chunk))
    first = [min(13, b["max_attempts"]) for b in blocks]
    oc, d_na, _, _ = eng.play_h2h(42, pair, order, st(blocks, "seat1"), st(blocks, "seat2"),
                                  [0] * len(blocks), first)
    prog = eng.h2h_resolve(d_na, oc, req, np.zeros((len(blocks), 5), dtype=np.int32))
    for b, p in zip(blocks, prog):
        assert p.tolist() == [b["after_chunk13"][k_] for k_ in keys]
    # then everything up to max_attempts; the resolve kernel applies the early stop
    rest = [b["max_attempts"] - int(p[0]) for b, p in zip(blocks, prog)]
    oc, d_na, _, _ = eng.play_h2h(42, pair, order, st(blocks, "seat1"), st(blocks, "seat2"),
                                  prog[:, 0], rest)
    prog = eng.h2h_resolve(d_na, oc, req, prog)
    for b, p in zip(blocks, prog):
        assert p.tolist() == [b["final"][k_] for k_ in keys]


def test_h2h_many_blocks_vs_oracle(eng, golden_dir):
    table = np.load(golden_dir / "games_full_0_2.npz")["strategies"]
    rng = np.random.Generator(np.random.PCG64DXSM(5))
    nb = 60
    a_idx, b_idx = rng.integers(0, 5160, size=nb), rng.integers(0, 5160, size=nb)
    pair = rng.integers(0, 10_000, size=nb)
    order = rng.integers(0, 2, size=nb)
    req = rng.integers(5, 40, size=nb)
    n_att = rng.integers(1, 70, size=nb)
    att0 = rng.integers(0, 1000, size=nb)
    oc, d_na, rows, totals = eng.play_h2h(77, pair, order, table[a_idx], table[b_idx], att0,
                                          n_att, want_rows=True)
    start = np.zeros((nb, 5), dtype=np.int32)
    start[:, 0] = att0
    prog = eng.h2h_resolve(d_na, oc, req, start)
    oc = oc.cpu().numpy()
    off = 0
    for i in range(nb):
        p, want_oc = fo.play_h2h_block(77, int(pair[i]), int(order[i]), table[a_idx[i]],
                                       table[b_idx[i]], n_completed_required=int(req[i]),
                                       max_attempts=int(att0[i] + n_att[i]),
                                       chunk_games=int(n_att[i]), progress=start[i])
        assert prog[i].tolist() == p.tolist(), i
        assert oc[off: off + len(want_oc)].tolist() == want_oc.tolist()
        off += int(n_att[i])
    hrows = rows.cpu().numpy().view(fo.row_dtype(2)).reshape(-1)
    assert int(hrows["game_seed"][0]) == fo.coordinate_seed(
        P_H2H_GAME, root_seed=77, k=2, pair_id=int(pair[0]), order=int(order[0]),
        game_index=int(att0[0]))
    assert totals.cpu().numpy()[0] == n_att.sum()


@pytest.mark.parametrize("k", [2, 4])
def test_full_size_cell_properties(eng, golden_dir, k):
    """BASELINE.json configs[1] at full size (4,300 shuffles of the 5,160 grid): size-independent
    invariants on the whole cell, bit-exact per-batch tallies against the oracle on three of
    the 100 deterministic batches, and the sum of the batch slots equal to a single-slot run."""
    table = np.load(golden_dir / "games_full_0_2.npz")["strategies"]
    n, nsh, spb = len(table), 4300, 43
    res = eng.play_tournament(42, k, 0, nsh, table, shuffles_per_slot=spb)
    t = res.tallies.cpu().numpy()
    tot = res.totals.cpu().numpy()
    games = nsh * (n // k)
    assert t.shape == (100, n, 26) and tot[0] == games and tot[1] + tot[2] == games and tot[7] == 0
    assert (t[:, :, 1] == spb).all()                              # one seat per strategy per shuffle
    assert (t[:, :, 2] + t[:, :, 3] == t[:, :, 1]).all()          # completed + safety = attempted
    assert t[:, :, 0].sum() == tot[1] and t[:, :, 3].sum() == k * tot[2]
    assert tot[8:8 + k].sum() == tot[1] and not tot[8 + k:].any()  # seat wins
    assert (t[:, :, 14] == 0).all() and (t[:, :, 25] == 0).all()  # winner_hit_max_rounds
    assert (t[:, :, 4] >= 10_000 * t[:, :, 0]).all()              # every winner reached the target
    assert (t[:, :, 15] >= t[:, :, 4] * 10_000).all()             # sum v^2 >= 10,000 * sum v
    for slot in (0, 57, 99):
        want, _, _ = fo.play_tournament(42, k, slot * spb, spb, table, n_threads=8)
        assert np.array_equal(t[slot], want[0]), slot
    one = eng.play_tournament(42, k, 0, nsh, table)
    assert np.array_equal(one.tallies.cpu().numpy()[0], t.sum(axis=0))
    assert np.array_equal(one.totals.cpu().numpy(), tot)


def test_run_tournament_end_to_end(eng, golden_dir, tmp_path):
    """run_tournament(): checkpoint payload, {k}p_metrics.parquet and per-shuffle row shards for
    the fast grid, seed 42, k=2 — the cell the reference's `farkle run` digest was taken on."""
    import pickle

    import pyarrow.parquet as pq

    from farkle_ii_b200 import run_tournament as frt
    from farkle_ii_b200.strategies import generate_strategy_grid

    z = np.load(golden_dir / "fast42.npz")
    strategies, _ = generate_strategy_grid(
        score_thresholds=[250, 300, 350, 400], smart_five_opts=[True], smart_one_opts=[True],
        consider_score_opts=[True], consider_dice_opts=[True], auto_hot_dice_opts=[True],
        run_up_score_opts=[True])
    cfg = frt.TournamentConfig(n_players=2, num_shuffles=600, deterministic_batch_size=30)
    ckpt = tmp_path / "2p_checkpoint.pkl"
    frt.run_tournament(config=cfg, global_seed=42, checkpoint_path=ckpt, collect_metrics=True,
                       num_shuffles=600, strategies=strategies, checkpoint_metadata={"seed": 42})
    payload = pickle.loads(ckpt.read_bytes())
    wins = payload["win_totals"]
    assert isinstance(wins, frt.OutcomeCounter) and payload["meta"]["seed"] == 42
    assert payload["meta"]["completed_shuffle_indices"] == list(range(600))
    assert payload["meta"]["completed_process_block_indices"] == list(range(1, 21))
    assert (wins.games_attempted, wins.games_completed, wins.games_safety_limit) == (24000, 23801, 199)
    assert [wins[i] for i in (42, 46, 51, 37, 25)] == [330, 322, 407, 386, 140]
    want = z["tallies"]
    for sid in range(80):
        assert wins[sid] == want[sid, 0] and wins.attempted_exposures[sid] == want[sid, 1]
        assert wins.completed_exposures[sid] == want[sid, 2]
        assert wins.safety_limit_exposures[sid] == want[sid, 3]
        for m, label in enumerate(frt.METRIC_LABELS):
            assert payload["metric_sums"][label].get(sid, 0.0) == float(want[sid, 4 + m])
            assert payload["metric_square_sums"][label].get(sid, 0.0) == float(want[sid, 15 + m])
    metrics = pq.read_table(tmp_path / "2p_metrics.parquet").to_pandas()
    assert len(metrics) == 11 * int((want[:, 0] > 0).sum())
    ws = metrics[metrics.metric == "winning_score"].set_index("strategy")["sum"]
    assert ws.sum() == 252_520_900
    # rows mode on a slice: one shard + one manifest line per shuffle, resumable
    cfg2 = frt.TournamentConfig(n_players=2, num_shuffles=4, deterministic_batch_size=2)
    row_dir = tmp_path / "rows"
    frt.run_tournament(config=cfg2, global_seed=42, checkpoint_path=tmp_path / "c2.pkl",
                       row_output_directory=row_dir, num_shuffles=4, strategies=strategies)
    shards = sorted(row_dir.glob("rows_42_2p_*.parquet"))
    assert len(shards) == 4 and len((row_dir / "manifest.jsonl").read_text().splitlines()) == 4
    first = pq.read_table(shards[0]).to_pylist()[0]
    assert (first["P1_strategy"], first["P2_strategy"]) == (42, 9)        # SURVEY.md §8c first row
    assert (first["winner_seat"], first["P1_score"], first["P2_score"], first["n_rounds"]) == ("P1", 10800, 6900, 17)
    assert (first["P1_rolls"], first["P1_farkles"], first["P1_highest_turn"], first["P1_hot_dice"]) == (36, 6, 2750, 5)
    assert (first["P2_rolls"], first["P2_farkles"], first["P2_smart_five_uses"]) == (52, 8, 20)
    assert first["shuffle_seed"] == 1_998_876_487 and first["game_seed"] == 1_433_242_307
    frt.run_tournament(config=cfg2, global_seed=42, checkpoint_path=tmp_path / "c2.pkl",
                       row_output_directory=row_dir, num_shuffles=4, strategies=strategies)
    assert len((row_dir / "manifest.jsonl").read_text().splitlines()) == 4  # resume skipped all


def test_longest_first_scheduling_and_tiny_launches(eng):
    """Tables made of never-banking strategies (the longest-first list covers every game), mixed
    tables, and launches smaller than a warp: scheduling must not change any result."""
    never = [(300, 0, 0x04 | 0x08 | 0x10), (0, 0, 0x08), (250, 0, 0x04 | 0x08 | 0x10 | 0x20)]
    normal = [(300, 2, 0x04 | 0x08), (500, 1, 0x01 | 0x04 | 0x08 | 0x20 | 0x80), (350, 3, 0x04 | 0x40)]
    for entries, k, nsh, mr in ((never[:2], 2, 3, 40), (never + normal, 2, 9, 25), (never + normal, 3, 7, 30),
                                (never + normal, 6, 5, 200), (normal[:1], 1, 4, 200)):
        table = np.array(entries, dtype=fo.STRATEGY_DTYPE)
        res = eng.play_tournament(77, k, 2, nsh, table, max_rounds=mr, want_rows=True, want_game_seeds=True)
        want_t, want_tot, want_rows = fo.play_tournament(77, k, 2, nsh, table, max_rounds=mr, want_rows=True,
                                                         want_game_seeds=True)
        assert np.array_equal(res.tallies.cpu().numpy(), want_t), (k, nsh)
        assert np.array_equal(res.totals.cpu().numpy(), want_tot), (k, nsh)
        assert res.rows_numpy().tobytes() == want_rows.tobytes()


def test_large_grid_global_memory_permutation(eng):
    """More strategies than the shared-memory Fisher-Yates holds (fallback permute_kernel)."""
    rng = np.random.Generator(np.random.PCG64DXSM(5))
    n = 120_000
    table = np.zeros(n, dtype=fo.STRATEGY_DTYPE)
    table["score_threshold"] = rng.integers(4, 20, size=n) * 50
    table["dice_threshold"] = rng.integers(1, 5, size=n)
    table["flags"] = 0x04 | 0x08 | np.where(rng.integers(0, 2, size=n) == 1, 0x20, 0)
    perm = eng.permute_shuffles(9, 4, 5, 2, n)
    for j in range(2):
        assert np.array_equal(perm[j], fo.permutation(9, 4, 5 + j, n))
    res = eng.play_tournament(9, 4, 5, 2, table)
    want_t, want_tot, _ = fo.play_tournament(9, 4, 5, 2, table, n_threads=8)
    assert np.array_equal(res.tallies.cpu().numpy(), want_t)
    assert np.array_equal(res.totals.cpu().numpy(), want_tot)


@pytest.mark.parametrize("spb", [0, 43])
def test_host_call_streams_rows_in_chunks(eng, golden_dir, spb):
    """fb_run_tournament_host in rows mode cuts the range into chunks and copies the rows of one
    chunk to the host under the kernels of the next: same rows, tallies and totals as one launch."""
    table = np.load(golden_dir / "games_full_0_2.npz")["strategies"]
    nsh = 500                                   # 1.29 M games: above the chunking threshold
    t, tot, rows = eng.run_tournament_host(7, 2, 11, nsh, table, shuffles_per_slot=spb, want_rows=True,
                                           want_game_seeds=True)
    one = eng.play_tournament(7, 2, 11, nsh, table, shuffles_per_slot=spb, want_rows=True,
                              want_game_seeds=True)
    assert rows.tobytes() == one.rows_numpy().tobytes()
    assert np.array_equal(rows["game_ordinal"], np.arange(len(rows), dtype=np.uint32))
    assert np.array_equal(t, one.tallies.cpu().numpy())
    assert np.array_equal(tot, one.totals.cpu().numpy())
    want_t, _, want_rows = fo.play_tournament(7, 2, 11 + 43 * 3, 43, table, want_rows=True,
                                              want_game_seeds=True, n_threads=8)
    gps = len(table) // 2
    raw = rows.view(np.uint8).reshape(len(rows), -1)[43 * 3 * gps:43 * 4 * gps].copy()  # keeps the padding bytes
    got = raw.reshape(-1).view(rows.dtype)
    got["game_ordinal"] -= 43 * 3 * gps
    assert got.tobytes() == want_rows.tobytes()
    if spb:
        assert np.array_equal(t[3], want_t[0])


@pytest.mark.parametrize("name,spb", [("fast_54_4", 2), ("fast_42_2", 5), ("full_0_5", 0)])
def test_seat_tallies(eng, golden_dir, name, spb):
    """Per (batch, strategy, seat) wins / exposures / completed / safety-limit counts from the
    gather pass equal the counts the reference's seat analysis derives from the rows
    (analysis/seat_analysis.py:166-229; the rows themselves are golden)."""
    from farkle_ii_b200 import run_tournament as frt

    z = np.load(golden_dir / f"games_{name}.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    res = eng.play_tournament(root, k, sh0, nsh, z["strategies"], shuffles_per_slot=spb,
                              want_seat_tallies=True)
    seat = res.seat_tallies.cpu().numpy()
    n = len(z["strategies"])
    gps = n // k
    batch = (np.arange(len(z["rows"])) // gps) // (spb if spb else nsh)
    want = frt.seat_counts_from_rows(z["rows"], batch)
    got = {(int(b), int(s), int(q) + 1): seat[b, s, q].tolist()
           for b, s, q in zip(*np.nonzero(seat[..., 1]))}
    assert got == want
    assert (seat[..., 1] == seat[..., 2] + seat[..., 3]).all()
    assert np.array_equal(seat[..., 0].sum(axis=2), res.tallies.cpu().numpy()[..., 0])
    tbl = frt.seat_counts_table(seat, np.arange(n), root_seed=root, k=k)
    assert tbl.num_rows == len(want) and tbl.column("raw_exposures").to_pylist() == [
        want[key][1] for key in sorted(want)]


def test_full_size_cell_bit_exact(eng, golden_dir):
    """The whole BASELINE cell — full grid, k=2, root 43, all 4,300 shuffles = 11,094,000 games —
    bit for bit against the oracle: every per-batch tally slot and the launch totals."""
    import os

    table = np.load(golden_dir / "games_full_0_2.npz")["strategies"]
    res = eng.play_tournament(43, 2, 0, 4300, table, shuffles_per_slot=43)
    want_t, want_tot, _ = fo.play_tournament(43, 2, 0, 4300, table, shuffles_per_slot=43,
                                             n_threads=os.cpu_count() or 8)
    assert np.array_equal(res.tallies.cpu().numpy(), want_t)
    assert np.array_equal(res.totals.cpu().numpy(), want_tot)


def test_random_tables_fuzz(eng):
    """Random strategy tables (thresholds that are not multiples of 50, dice thresholds -1..6, every
    legal flag combination incl. "considers nothing"), random k / target / safety limit: rows,
    tallies and totals against the oracle."""
    rng = np.random.Generator(np.random.PCG64DXSM(20260118))
    for case in range(24):
        k = int(rng.choice([1, 2, 2, 3, 4, 5, 6, 8, 12]))
        n = k * int(rng.integers(2, 40))
        table = np.zeros(n, dtype=fo.STRATEGY_DTYPE)
        table["score_threshold"] = rng.integers(0, 1600, size=n) if case % 2 else rng.integers(0, 30, size=n) * 50
        table["dice_threshold"] = rng.integers(-1, 7, size=n)
        flags = rng.integers(0, 256, size=n)
        flags &= np.where(flags & 0x01, 0xFF, 0xFF & ~0x02)                      # smart_one needs smart_five
        flags &= np.where((flags & 0x0C) == 0x0C, 0xFF, 0xFF & ~0x10)            # require_both needs both
        table["flags"] = flags
        target = int(rng.choice([200, 1000, 2500, 10_000]))
        max_rounds = int(rng.choice([1, 3, 12, 200]))
        nsh = int(rng.integers(1, 60))
        root, sh0 = int(rng.integers(0, 2**62)), int(rng.integers(0, 2**40))
        res = eng.play_tournament(root, k, sh0, nsh, table, target_score=target, max_rounds=max_rounds,
                                  want_rows=True, want_game_seeds=True, shuffles_per_slot=int(rng.integers(0, 4)))
        want_t, want_tot, want_rows = fo.play_tournament(
            root, k, sh0, nsh, table, target_score=target, max_rounds=max_rounds, want_rows=True,
            want_game_seeds=True, shuffles_per_slot=0, n_threads=4)
        assert res.rows_numpy().tobytes() == want_rows.tobytes(), case
        assert np.array_equal(res.tallies.cpu().numpy().sum(axis=0), want_t[0]), case
        assert np.array_equal(res.totals.cpu().numpy(), want_tot), case


@pytest.mark.parametrize("name,with_ids", [("fast_54_4", False), ("fast_42_2", True), ("full_0_5", False),
                                           ("full_0_2", False)])
def test_first_seen_ordinals(eng, golden_dir, name, with_ids):
    """`first_seen` (key insertion order of the reference's counters): first win / exposure /
    completed / safety-limit exposure ordinal per id == a walk over the rows in game and seat order."""
    from oracle_engine import OracleEngine

    z = np.load(golden_dir / f"games_{name}.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    table = z["strategies"]
    ids = (np.arange(len(table), dtype=np.int32)[::-1] * 2 + 5).copy() if with_ids else None
    res = eng.play_tournament(root, k, sh0, nsh, table, strategy_ids=ids, want_first_seen=True, want_rows=True)
    want = OracleEngine._first_seen(res.rows_numpy(), len(table), k, ids)
    got = res.first_seen.cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got, want)
    assert (got[:, 1] >= 0).sum() == len(table)          # every strategy is seated in the first shuffle
    assert sorted(got[got[:, 1] >= 0, 1].tolist()) == list(range(len(table)))
    # slotted tallies and several gather chunks do not change it
    res2 = eng.play_tournament(root, k, sh0, nsh, table, strategy_ids=ids, want_first_seen=True,
                               shuffles_per_slot=2)
    assert np.array_equal(res2.first_seen.cpu().numpy(), want)
