"""Pin the CPU oracle (oracle/farkle_oracle.c) to the reference.

Every fixture under tests/golden/ was produced by tests/golden/make_golden.py
importing the unmodified reference; the known-answer values typed in below are
the reference's own (file:line cited next to each).
"""

from __future__ import annotations

import json

import numpy as np
import pytest

import oracle as fo

P_PLAYER, P_SHUFFLE, P_PERM, P_GAME, P_TPLAYER = 10, 100, 101, 102, 103


def _load(golden_dir, name):
    return np.load(golden_dir / name)


# --------------------------------------------------------------------------- RNG
def test_seedseq_words(golden_dir):
    data = json.loads((golden_dir / "rng.json").read_text())
    for case in data["seedseq"]:
        got = fo.seedseq_generate(case["entropy"], 8)
        assert got.tolist() == case["words"]


def test_reference_seedsequence_kat():
    # /root/reference/tests/unit/utils/test_random_utils.py:32-40,73-78
    mask = 2**32 - 1
    for root, k, sh, g in ((32, 2, 194, 18), (32, 2, 4052, 4)):
        entropy = [1, 102]
        for v in (root, k, sh, 0, 0, g, 0, 0):
            entropy.extend((v & mask, v >> 32))
        assert int(fo.seedseq_generate(entropy, 1)[0]) == 2_963_478_802


def test_survey_fingerprints():
    # SURVEY.md §8c, produced by running the reference
    assert fo.coordinate_seed(P_GAME, root_seed=11, k=2, shuffle_index=3, game_index=1,
                              as_u32=True) == 3_264_471_624
    assert fo.coordinate_seed(P_GAME, root_seed=54, k=4, shuffle_index=100, game_index=7,
                              as_u32=True) == 920_640_453
    assert fo.coordinate_seed(P_SHUFFLE, root_seed=42, k=2, shuffle_index=0,
                              as_u32=True) == 1_998_876_487
    assert fo.coordinate_seed(P_SHUFFLE, root_seed=42, k=2, shuffle_index=1,
                              as_u32=True) == 2_468_994_662
    assert fo.coordinate_seed(P_GAME, root_seed=42, k=2, shuffle_index=0, game_index=0,
                              as_u32=True) == 1_433_242_307


def test_coordinate_streams(golden_dir):
    data = json.loads((golden_dir / "rng.json").read_text())
    pattern = data["n_dice_pattern"]
    for case in data["coords"]:
        c = dict(case["coord"])
        purpose = c.pop("purpose")
        assert fo.coordinate_entropy(purpose, **c).tolist() == case["entropy"]
        assert fo.coordinate_seed(purpose, as_u32=True, **c) == case["seed_u32"]
        assert fo.coordinate_seed(purpose, as_u32=False, **c) == case["seed_u64"]
        si = fo.seed_stream(purpose, **c)
        state, inc = int(case["state"]), int(case["inc"])
        assert (int(si[0]) << 64) | int(si[1]) == state
        assert (int(si[2]) << 64) | int(si[3]) == inc
        faces = fo.roll_dice_state(si, pattern)
        for r, n in enumerate(pattern):
            assert faces[r, :n].tolist() == case["dice"][r]
            assert not faces[r, n:].any()


def test_first_rolls_kat():
    # PLAYER(10) root 42, k 2, game 0, seat 0
    si = fo.seed_stream(P_PLAYER, root_seed=42, k=2, game_index=0, seat_index=0)
    faces = fo.roll_dice_state(si, [6] * 6)
    assert faces.tolist() == [[2, 2, 6, 4, 3, 5], [1, 6, 4, 4, 2, 3], [5, 3, 6, 2, 2, 2],
                              [2, 1, 4, 5, 1, 5], [1, 2, 2, 2, 3, 6], [6, 6, 5, 5, 2, 5]]


def test_lemire_rejection_branch(golden_dir):
    data = json.loads((golden_dir / "rng.json").read_text())
    assert len(data["rejection"]) >= 8
    for case in data["rejection"]:
        state, inc = int(case["state"]), int(case["inc"])
        m = 2**64 - 1
        si = np.array([state >> 64, state & m, inc >> 64, inc & m], dtype=np.uint64)
        faces = fo.roll_dice_state(si, case["n_dice"], has32=case.get("has32", 0),
                                   saved=case.get("saved", 0))
        for r, n in enumerate(case["n_dice"]):
            assert faces[r, :n].tolist() == case["dice"][r]


def test_permutations(golden_dir):
    perms = _load(golden_dir, "perm.npz")
    for key in perms.files:
        root, k, sh, n = (int(x) for x in key.split("_"))
        assert np.array_equal(fo.permutation(root, k, sh, n), perms[key]), key


# ----------------------------------------------------------------------- scoring
def test_score_table_923(golden_dir):
    tab = _load(golden_dir, "scoring.npz")["table"]
    assert tab.shape == (923, 10)
    for row in tab:
        assert fo.evaluate_counts(row[:6]) == tuple(int(x) for x in row[6:])


def test_reference_golden_rolls(golden_dir):
    # /root/reference/tests/data/test_farkle_scores_data.csv via tests/unit/game/test_scoring.py:184-194
    rolls = _load(golden_dir, "scoring.npz")["csv_rolls"]
    assert len(rolls) == 153
    plain = np.zeros(1, dtype=fo.STRATEGY_DTYPE)  # smart flags off
    for row in rolls:
        faces = [int(f) for f in row[:6] if f]
        counts = [faces.count(f) for f in range(1, 7)]
        score, used, sf, so = fo.evaluate_counts(counts)
        assert (score, used, len(faces) - used, sf, so) == tuple(int(x) for x in row[6:])
        assert fo.default_score(faces, 0, plain)[:3] == (score, used, len(faces) - used)


def _strategy(st, dt, flags):
    s = np.zeros(1, dtype=fo.STRATEGY_DTYPE)
    s["score_threshold"], s["dice_threshold"], s["flags"] = st, dt, flags
    return s


def test_discard_cases(golden_dir):
    disc = _load(golden_dir, "scoring.npz")["discards"]
    assert len(disc) == 40
    for row in disc:
        counts = row[:6]
        faces = [f + 1 for f in range(6) for _ in range(int(counts[f]))]
        ts, st, dt, cs, cd, rb, sf, so, d5, d1 = (int(x) for x in row[6:])
        flags = sf | so << 1 | cs << 2 | cd << 3 | rb << 4 | 0x80
        out = fo.default_score(faces, ts, _strategy(st, dt, flags))
        assert out[3:] == (d5, d1)


def test_default_score_sweep(golden_dir):
    z = _load(golden_dir, "scoring.npz")
    for row, want in zip(z["sweep_in"], z["sweep_out"], strict=True):
        faces = [int(f) for f in row[:6] if f]
        ts, st, dt, flags = (int(x) for x in row[6:])
        assert fo.default_score(faces, ts, _strategy(st, dt, flags)) == tuple(int(x) for x in want)


# ------------------------------------------------------------------- whole games
GAME_FIXTURES = ["fast_42_2", "fast_54_4", "fast_54_5", "full_0_2", "full_42_4", "full_0_5",
                 "full_42_6", "full_102_12", "full_102_3"]


def assert_rows_equal(got, want, k):
    for name in ("game_seed", "n_rounds", "winner_seat", "flags"):
        assert np.array_equal(got[name], want[name]), name
    for name in fo.SEAT_DTYPE.names:
        assert np.array_equal(got["seats"][name], want["seats"][name]), name


@pytest.mark.parametrize("name", GAME_FIXTURES)
def test_tournament_rows_and_tallies(golden_dir, name):
    z = _load(golden_dir, f"games_{name}.npz")
    root, k, sh0, nsh = (int(x) for x in z["meta"])
    tallies, totals, rows = fo.play_tournament(root, k, sh0, nsh, z["strategies"], want_rows=True,
                                               want_game_seeds=True, n_threads=2)
    assert_rows_equal(rows, z["rows"], k)
    assert np.array_equal(tallies[0], z["tallies"])
    assert totals[0] == len(rows) and totals[1] + totals[2] == totals[0]
    for i, sh in enumerate(range(sh0, sh0 + nsh)):
        assert fo.coordinate_seed(P_SHUFFLE, root_seed=root, k=k, shuffle_index=sh,
                                  as_u32=True) == int(z["shuffle_seeds"][i])


# /root/reference/tests/integration/test_raw_simulation_oracle.py:45-58
EXPECTED_ROWS = {
    (11, 2, 0, 0): ([0, 2], "safety_limit", None, 0, 0, [0, 0]),
    (11, 2, 0, 1): ([1, 3], "completed", 1, 2, 5, [1950, 1100]),
    (11, 2, 1, 0): ([2, 1], "completed", 2, 2, 4, [500, 0]),
    (11, 2, 1, 1): ([0, 3], "completed", 0, 1, 2, [600, 0]),
    (11, 4, 0, 0): ([0, 1, 2, 3], "completed", 2, 1, 4, [700, 0, 800, 0]),
    (11, 4, 1, 0): ([3, 2, 1, 0], "completed", 3, 1, 5, [3050, 2900, 0, 0]),
    (22, 2, 0, 0): ([3, 0], "completed", 3, 1, 2, [600, 0]),
    (22, 2, 0, 1): ([1, 2], "completed", 2, 1, 2, [500, 1100]),
    (22, 2, 1, 0): ([2, 0], "completed", 2, 1, 2, [950, 0]),
    (22, 2, 1, 1): ([3, 1], "completed", 3, 1, 3, [750, 550]),
    (22, 4, 0, 0): ([1, 2, 0, 3], "completed", 2, 1, 5, [0, 700, 0, 0]),
    (22, 4, 1, 0): ([0, 2, 1, 3], "completed", 1, 1, 4, [700, 0, 1100, 0]),
}
ORACLE_OVERRIDES = {(11, 2): [(0, 0, 0)]}  # tests/helpers/raw_simulation_oracle.py:57-77


def check_expected_rows(play):
    """Shared by the oracle test and the GPU parity test: play(root, k, overrides)."""
    seen = 0
    for root in (11, 22):
        for k in (2, 4):
            rows = play(root, k, ORACLE_OVERRIDES.get((root, k), []))
            gps = 4 // k
            for i, row in enumerate(rows):
                want = EXPECTED_ROWS[(root, k, i // gps, i % gps)]
                strategies, status, winner_strategy, rounds, turns, scores = want
                assert row["seats"]["strategy"].tolist() == strategies
                safety = bool(row["flags"] & 1)
                assert ("safety_limit" if safety else "completed") == status
                got_winner = None if safety else int(row["seats"]["strategy"][row["winner_seat"]])
                assert got_winner == winner_strategy
                assert int(row["n_rounds"]) == rounds
                assert int(row["seats"]["n_turns"].sum()) == turns
                assert row["seats"]["score"].tolist() == scores
                seen += 1
    assert seen == 12


def test_reference_raw_oracle_12_games(golden_dir):
    z = _load(golden_dir, "oracle12.npz")

    def play(root, k, overrides):
        tallies, _totals, rows = fo.play_tournament(root, k, 0, 2, z["strategies"],
                                                    target_score=100, overrides=overrides,
                                                    want_rows=True, want_game_seeds=True)
        assert_rows_equal(rows, z[f"rows_{root}_{k}"], k)
        assert np.array_equal(tallies[0], z[f"tallies_{root}_{k}"])
        return rows

    check_expected_rows(play)


def test_public_helper_kat(golden_dir):
    # /root/reference/tests/unit/simulation/test_simulation.py:184-198
    z = _load(golden_dir, "helpers.npz")
    coords = np.array([[P_PLAYER, 123, 3, 0, 0, 0, g] for g in range(10)], dtype=np.uint64)
    strat = np.tile(z["strategies"], (10, 1))
    rows, totals = fo.play_games(coords, 3, strat, target_score=5000)
    winners = np.bincount(rows["winner_seat"], minlength=3).tolist()
    assert winners == [2, 6, 2]  # {"P2": 6, "P1": 2, "P3": 2}
    want = z["rows"].copy()
    want["game_seed"] = 0  # fo.play_games does not compute the helper fingerprint
    assert_rows_equal(rows, want, 3)
    for g in range(10):  # spawn_seeds: INDEXED_SEED(1) fingerprints, utils/random.py:275-295
        assert fo.coordinate_seed(1, root_seed=123, game_index=g, as_u32=True) == int(
            z["game_seeds"][g])


def test_h2h_blocks(golden_dir):
    blocks = json.loads((golden_dir / "h2h.json").read_text())
    assert len(blocks) == 5
    keys = ("games_attempted", "games_completed", "games_safety_limit", "wins_seat1", "wins_seat2")
    saw_safety = False
    for b in blocks:
        s1, s2 = _strategy(*b["seat1"]), _strategy(*b["seat2"])
        kw = dict(n_completed_required=b["n_completed_required"], max_attempts=b["max_attempts"])
        p1, oc1 = fo.play_h2h_block(b["root_seed"], b["pair_id"], b["order"], s1, s2,
                                    chunk_games=13, **kw)
        assert p1.tolist() == [b["after_chunk13"][k] for k in keys]
        p2, oc2 = fo.play_h2h_block(b["root_seed"], b["pair_id"], b["order"], s1, s2,
                                    chunk_games=5000, progress=p1, **kw)
        assert p2.tolist() == [b["final"][k] for k in keys]
        assert len(oc1) + len(oc2) == p2[0]
        saw_safety |= p2[2] > 0
    assert saw_safety


def test_fast_grid_seed42_full_cell(golden_dir):
    # reference `farkle run`, fast grid, seed 42, k=2, 600 shuffles
    z = _load(golden_dir, "fast42.npz")
    tallies, totals, _ = fo.play_tournament(42, 2, 0, 600, z["strategies"], n_threads=4)
    assert np.array_equal(tallies[0], z["tallies"])
    assert totals[:3].tolist() == [24000, 23801, 199]
    assert tallies[0][[42, 46, 51, 37, 25], 0].tolist() == [330, 322, 407, 386, 140]
    assert int(tallies[0][:, 4].sum()) == 252_520_900   # sum winning_score
    assert int(tallies[0][:, 5].sum()) == 474_671       # sum n_rounds
    assert int(tallies[0][:, 7].sum()) == 1_180_523     # sum winner_rolls


def test_rejected_half_fixture(golden_dir):
    """tests/golden/rejects.json (games whose dice meet a half the Lemire test rejects; the GPU suite
    plays exactly these): the oracle finds the rejected halves where the fixture says, and nowhere
    else in those shuffles."""
    import json

    from farkle_ii_b200.strategies import generate_strategy_grid, pack_strategies

    table = pack_strategies(generate_strategy_grid()[0])
    games = json.loads((golden_dir / "rejects.json").read_text())["games"]
    assert len(games) == 18 and {g["k"] for g in games} == {2, 3, 4}
    for g in games[::4]:
        hits = fo.scan_rejected_halves(g["root"], g["k"], g["shuffle"], 1, table)
        assert hits.tolist() == [[g["shuffle"], g["game"], g["rejects"]]]
    # and a shuffle without one
    assert len(fo.scan_rejected_halves(42, 2, 0, 1, table)) == 0

