#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden.py

It imports the unmodified reference from ``/root/reference/src`` and freezes
the outputs of its own functions for the hot path:

* ``rng.json``       SeedSequence words, coordinate_seed fingerprints, PCG64DXSM
                     (state, inc) of coordinate_rng, dice of Generator.integers
* ``perm.npz``       Generator.permutation of SHUFFLE_PERMUTATION streams
* ``scoring.npz``    the 923-entry SCORE_TABLE, the reference's 153 golden rolls
                     (tests/data/test_farkle_scores_data.csv), its 40 discard
                     cases, and a 40k-case default_score sweep
* ``games_*.npz``    whole-game rows + tallies of _play_one_shuffle for several
                     (grid, root, k, shuffle) cells, compact-row encoded
* ``oracle12.npz``   the 12 games of tests/integration/test_raw_simulation_oracle.py
* ``helpers.npz``    simulate_many_games KAT rows (tests/unit/simulation/test_simulation.py:184-198)
* ``h2h.json``       _simulate_block_from_manifest progress for a few blocks
* ``fast42.npz``     full fast grid, seed 42, k=2, 600 shuffles: tallies[80][26]

No reference source is copied; only values it computes.
"""

from __future__ import annotations

import csv
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))

import numpy as np  # noqa: E402

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from farkle.game import scoring as ref_scoring  # noqa: E402
from farkle.simulation import run_tournament as rt  # noqa: E402
from farkle.simulation import simulation as ref_sim  # noqa: E402
from farkle.simulation.game_profile import GameProfile, TournamentMaxRoundsOverride  # noqa: E402
from farkle.simulation.strategies import (  # noqa: E402
    FavorDiceOrScore,
    ThresholdStrategy,
    build_strategy_manifest,
)
from farkle.utils import random as ur  # noqa: E402

from oracle import STRATEGY_DTYPE, TALLY_WIDTH, row_dtype  # noqa: E402  (layout only)

SF = dict(smart_five=1, smart_one=2, consider_score=4, consider_dice=8, require_both=16,
          auto_hot_dice=32, run_up_score=64)


def pack_strategy(s: ThresholdStrategy) -> tuple[int, int, int]:
    flags = 0
    for name, bit in SF.items():
        if getattr(s, name):
            flags |= bit
    if s.favor_dice_or_score is FavorDiceOrScore.SCORE:
        flags |= 128
    return int(s.score_threshold), int(s.dice_threshold), flags


def pack_strategies(strats) -> np.ndarray:
    return np.array([pack_strategy(s) for s in strats], dtype=STRATEGY_DTYPE)


FAST_GRID = dict(score_thresholds=[250, 300, 350, 400], dice_thresholds=None,
                 smart_five_opts=[True], smart_one_opts=[True], consider_score_opts=[True],
                 consider_dice_opts=[True], auto_hot_dice_opts=[True], run_up_score_opts=[True])
TINY_GRID = dict(score_thresholds=[500], dice_thresholds=[2], smart_five_opts=[False],
                 smart_one_opts=[False], consider_score_opts=[True], consider_dice_opts=[True],
                 auto_hot_dice_opts=[False, True], run_up_score_opts=[False])


def grid(kind: str):
    if kind == "fast":
        return ref_sim.generate_strategy_grid(**FAST_GRID)[0]
    if kind == "tiny":
        return ref_sim.generate_strategy_grid(**TINY_GRID)[0]
    return ref_sim.generate_strategy_grid()[0]


# --------------------------------------------------------------------------- rows
def rows_to_compact(rows, k: int) -> np.ndarray:
    out = np.zeros(len(rows), dtype=row_dtype(k))
    for i, r in enumerate(rows):
        out["game_seed"][i] = r["game_seed"]
        out["game_ordinal"][i] = i
        out["n_rounds"][i] = r["n_rounds"]
        safety = r["termination_status"] == "safety_limit"
        out["winner_seat"][i] = 0xFF if safety else int(r["winner_seat"][1:]) - 1
        out["flags"][i] = 1 if safety else 0
        for s in range(k):
            p = f"P{s + 1}_"
            seat = out["seats"][i, s]
            seat["score"] = r[p + "score"]
            seat["strategy"] = r[p + "strategy"]
            seat["highest_turn"] = r[p + "highest_turn"]
            seat["farkles"] = r[p + "farkles"]
            seat["rolls"] = r[p + "rolls"]
            seat["n_turns"] = r[p + "n_turns"]
            seat["hot_dice"] = r[p + "hot_dice"]
            seat["smart_five_uses"] = r[p + "smart_five_uses"]
            seat["n_smart_five_dice"] = r[p + "n_smart_five_dice"]
            seat["smart_one_uses"] = r[p + "smart_one_uses"]
            seat["n_smart_one_dice"] = r[p + "n_smart_one_dice"]
            assert r[p + "hit_max_rounds"] == safety
    return out


def tallies_from(wins, sums, sqs, n_ids: int) -> np.ndarray:
    t = np.zeros((n_ids, TALLY_WIDTH), dtype=np.int64)
    for sid, v in wins.items():
        t[sid, 0] = v
    for sid, v in wins.attempted_exposures.items():
        t[sid, 1] = v
    for sid, v in wins.completed_exposures.items():
        t[sid, 2] = v
    for sid, v in wins.safety_limit_exposures.items():
        t[sid, 3] = v
    for i, label in enumerate(rt.METRIC_LABELS):
        for sid, v in sums[label].items():
            assert float(v).is_integer()
            t[sid, 4 + i] = int(v)
        for sid, v in sqs[label].items():
            assert float(v).is_integer()
            t[sid, 15 + i] = int(v)
    return t


def play_shuffles(kind: str, root: int, k: int, shuffles, *, profile=None, collect_rows=True):
    strats = grid(kind)
    cfg = rt.TournamentConfig(n_players=k, n_strategies=len(strats))
    rt._init_worker(strats, cfg, profile)
    all_rows, seeds = [], []
    tl = np.zeros((len(strats), TALLY_WIDTH), dtype=np.int64)
    for sh in shuffles:
        sseed = ur.coordinate_seed(ur.RandomPurpose.TOURNAMENT_SHUFFLE, root_seed=root, k=k,
                                   shuffle_index=sh, dtype=np.uint32)
        w, s, q, rows = rt._play_one_shuffle(rt.ShuffleTask(root, k, sh, sseed, 0),
                                             collect_rows=collect_rows)
        tl += tallies_from(w, s, q, len(strats))
        all_rows.extend(rows)
        seeds.append(sseed)
    return strats, all_rows, tl, seeds


def _fast42_chunk(args):
    lo, hi = args
    _, _, tl, _ = play_shuffles("fast", 42, 2, range(lo, hi), collect_rows=False)
    return tl


# --------------------------------------------------------------------------- main
def main() -> None:
    rng_out: dict = {}

    # --- SeedSequence words
    ss_cases = []
    for entropy in ([2, 103, 42, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [1, 2, 3], [7],
                    [1, 102, 32, 0, 2, 0, 194, 0, 0, 0, 0, 0, 18, 0, 0, 0, 0, 0],
                    [1, 102, 32, 0, 2, 0, 4052, 0, 0, 0, 0, 0, 4, 0, 0, 0, 0, 0],
                    list(range(100, 125))):
        ss = np.random.SeedSequence(entropy)
        ss_cases.append({"entropy": entropy,
                         "words": [int(x) for x in ss.generate_state(8, dtype=np.uint32)]})
    rng_out["seedseq"] = ss_cases

    # --- coordinate_seed / coordinate_rng states / dice
    coords = []
    gen = np.random.Generator(np.random.PCG64DXSM(12345))
    base = [
        dict(purpose=103, root_seed=42, k=2, shuffle_index=0, game_index=0, seat_index=0),
        dict(purpose=103, root_seed=42, k=2, shuffle_index=0, game_index=0, seat_index=1),
        dict(purpose=10, root_seed=42, k=2, game_index=0, seat_index=0),
        dict(purpose=101, root_seed=0, k=5, shuffle_index=7),
        dict(purpose=203, root_seed=11, k=2, pair_id=3, order=1, game_index=17, seat_index=1),
        dict(purpose=102, root_seed=11, k=2, shuffle_index=3, game_index=1),
        dict(purpose=102, root_seed=54, k=4, shuffle_index=100, game_index=7),
        dict(purpose=100, root_seed=42, k=2, shuffle_index=0),
        dict(purpose=100, root_seed=42, k=2, shuffle_index=1),
        dict(purpose=202, root_seed=2**63 + 5, k=2, pair_id=2**40 + 1, order=1, game_index=2**33),
    ]
    for _ in range(30):
        base.append(dict(purpose=int(gen.choice([10, 100, 101, 102, 103, 202, 203])),
                         root_seed=int(gen.integers(0, 2**63)), k=int(gen.integers(2, 13)),
                         shuffle_index=int(gen.integers(0, 2**34)),
                         pair_id=int(gen.integers(0, 50000)), order=int(gen.integers(0, 2)),
                         game_index=int(gen.integers(0, 5000)),
                         seat_index=int(gen.integers(0, 12))))
    n_dice_pattern = [6, 6, 3, 1, 2, 5, 4, 6, 1, 1, 2, 3, 6, 5, 5, 4]
    for c in base:
        kw = {k_: v for k_, v in c.items() if k_ != "purpose"}
        g = ur.coordinate_rng(c["purpose"], **kw)
        st = g.bit_generator.state
        dice = [[int(x) for x in g.integers(1, 7, size=n)] for n in n_dice_pattern]
        coords.append({
            "coord": c,
            "entropy": [int(x) for x in ur.coordinate_entropy(c["purpose"], **kw)],
            "seed_u32": ur.coordinate_seed(c["purpose"], dtype=np.uint32, **kw),
            "seed_u64": ur.coordinate_seed(c["purpose"], dtype=np.uint64, **kw),
            "state": str(st["state"]["state"]), "inc": str(st["state"]["inc"]),
            "dice": dice,
        })
    rng_out["n_dice_pattern"] = n_dice_pattern
    rng_out["coords"] = coords

    # --- crafted states that hit the Lemire rejection loop (low32(6*u) < 4)
    rej = []
    mult = 0xDA942042E4DD58B5
    m64 = 2**64 - 1
    for target_lo, hi_seed in ((715827883, 1), (1431655766, 99), (2863311531, 12345),
                               (3579139414, 2**60 + 3)):
        want = (0xABCDEF01 << 32) | target_lo  # a 64-bit output with that low half
        t = (want & -want).bit_length() - 1  # 2-adic valuation the mixed hi must share
        hi = hi_seed
        while True:
            h = hi ^ (hi >> 32)
            h = (h * mult) & m64
            h ^= h >> 48
            if h and ((h & -h).bit_length() - 1) == t:
                break
            hi += 1
        lo = ((want >> t) * pow(h >> t, -1, 2**64)) & m64
        assert lo & 1 and (h * lo) & m64 == want
        state = (hi << 64) | lo
        inc = (0x1234567 << 1) | 1
        bg = np.random.PCG64DXSM(0)
        bg.state = {"bit_generator": "PCG64DXSM", "state": {"state": state, "inc": inc},
                    "has_uint32": 0, "uinteger": 0}
        g = np.random.Generator(bg)
        pattern = [6, 2, 5, 6]
        dice = [[int(x) for x in g.integers(1, 7, size=n)] for n in pattern]
        rej.append({"state": str(state), "inc": str(inc), "n_dice": pattern, "dice": dice})
        # and once with the rejection landing on a buffered half
        bg.state = {"bit_generator": "PCG64DXSM", "state": {"state": state, "inc": inc},
                    "has_uint32": 1, "uinteger": target_lo}
        g = np.random.Generator(bg)
        dice = [[int(x) for x in g.integers(1, 7, size=n)] for n in pattern]
        rej.append({"state": str(state), "inc": str(inc), "has32": 1, "saved": target_lo,
                    "n_dice": pattern, "dice": dice})
    rng_out["rejection"] = rej
    (HERE / "rng.json").write_text(json.dumps(rng_out, indent=0))

    # --- permutations
    perm_out = {}
    for root, k, sh, n in ((42, 2, 0, 80), (42, 2, 1, 80), (0, 5, 0, 5160), (102, 12, 4299, 5160),
                           (11, 2, 0, 4), (11, 4, 1, 4), (7, 3, 9, 1), (7, 3, 9, 2), (7, 3, 9, 33),
                           (2**40, 6, 2**35, 257)):
        g = ur.coordinate_rng(ur.RandomPurpose.SHUFFLE_PERMUTATION, root_seed=root, k=k,
                              shuffle_index=sh)
        perm_out[f"{root}_{k}_{sh}_{n}"] = g.permutation(n).astype(np.int32)
    np.savez_compressed(HERE / "perm.npz", **perm_out)

    # --- scoring
    table = ref_scoring.SCORE_TABLE
    keys = sorted(table)
    tab = np.array([[*k_, table[k_][0], table[k_][1], table[k_][3], table[k_][4]] for k_ in keys],
                   dtype=np.int32)
    assert len(tab) == 923
    csv_rows = []
    with open(REF / "tests/data/test_farkle_scores_data.csv") as fh:
        for r in csv.DictReader(fh):
            faces = json.loads(r["Dice_Roll"])
            f6 = faces + [0] * (6 - len(faces))
            csv_rows.append(f6 + [int(r["Score"]), int(r["Used_Dice"]), int(r["Reroll_Dice"]),
                                  int(r["Single_Fives"]), int(r["Single_Ones"])])
    csv_rows = np.array(csv_rows, dtype=np.int32)
    assert len(csv_rows) == 153
    # discard CSV (unused by the reference's tests but valid data): re-evaluate through
    # the reference to freeze the *reference's* answer next to the file's expectation
    disc = []
    with open(REF / "tests/data/test_decide_smart_discards.csv") as fh:
        for r in csv.DictReader(fh):
            counts_map = eval(r["counts"])  # noqa: S307 - literal dict from the fixture
            counts = tuple(counts_map.get(f, 0) for f in range(1, 7))
            kw = dict(counts=counts, single_fives=int(r["single_fives"]),
                      single_ones=int(r["single_ones"]), raw_score=int(r["raw_score"]),
                      raw_used=int(r["raw_used"]), dice_roll_len=int(r["dice_len"]),
                      turn_score_pre=int(r["turn_score_pre"]),
                      score_threshold=int(r["score_threshold"]),
                      dice_threshold=int(r["dice_threshold"]),
                      consider_score=r["consider_score"] == "True",
                      consider_dice=r["consider_dice"] == "True",
                      require_both=r["require_both"] == "True",
                      smart_five=r["smart_five"] == "True", smart_one=r["smart_one"] == "True")
            got = ref_scoring.decide_smart_discards(**kw)
            disc.append([*counts, kw["turn_score_pre"], kw["score_threshold"],
                         kw["dice_threshold"], int(kw["consider_score"]), int(kw["consider_dice"]),
                         int(kw["require_both"]), int(kw["smart_five"]), int(kw["smart_one"]),
                         got[0], got[1]])
    disc = np.array(disc, dtype=np.int32)
    # default_score sweep
    g = np.random.Generator(np.random.PCG64DXSM(2024))
    sweep_in, sweep_out = [], []
    for _ in range(40000):
        n = int(g.integers(1, 7))
        faces = [int(x) for x in g.integers(1, 7, size=n)]
        if g.random() < 0.5:  # bias towards rolls with lone 1s / 5s
            faces[int(g.integers(0, n))] = int(g.choice([1, 5]))
        ts = int(g.integers(0, 40)) * 50
        sf = bool(g.integers(0, 2))
        so = bool(g.integers(0, 2)) and sf
        cs, cd = bool(g.integers(0, 2)), bool(g.integers(0, 2))
        rb = bool(g.integers(0, 2)) and cs and cd
        fav = FavorDiceOrScore.SCORE if g.integers(0, 2) else FavorDiceOrScore.DICE
        st_ = int(g.choice([199, 200, 250, 300, 350, 425, 500, 650, 800, 1000, 1350]))
        dt = int(g.integers(-1, 6))
        res = ref_scoring.default_score(faces, turn_score_pre=ts, smart_five=sf, smart_one=so,
                                        consider_score=cs, consider_dice=cd, require_both=rb,
                                        score_threshold=st_, dice_threshold=dt,
                                        favor_dice_or_score=fav, return_discards=True)
        flags = (sf * 1) | (so * 2) | (cs * 4) | (cd * 8) | (rb * 16) | (
            128 if fav is FavorDiceOrScore.SCORE else 0)
        sweep_in.append(faces + [0] * (6 - n) + [ts, st_, dt, flags])
        sweep_out.append(list(res))
    np.savez_compressed(HERE / "scoring.npz", table=tab, csv_rolls=csv_rows, discards=disc,
                        sweep_in=np.array(sweep_in, dtype=np.int32),
                        sweep_out=np.array(sweep_out, dtype=np.int32))

    # --- whole games
    for name, kind, root, k, shuffles in (
        ("fast_42_2", "fast", 42, 2, range(0, 12)),
        ("fast_54_4", "fast", 54, 4, range(3, 8)),
        ("fast_54_5", "fast", 54, 5, range(0, 4)),
        ("full_0_2", "full", 0, 2, [0]),
        ("full_42_4", "full", 42, 4, [17]),
        ("full_0_5", "full", 0, 5, [0]),
        ("full_42_6", "full", 42, 6, [4299]),
        ("full_102_12", "full", 102, 12, [1]),
        ("full_102_3", "full", 102, 3, [2]),
    ):
        strats, rows, tl, seeds = play_shuffles(kind, root, k, shuffles)
        np.savez_compressed(HERE / f"games_{name}.npz", strategies=pack_strategies(strats),
                            rows=rows_to_compact(rows, k), tallies=tl,
                            shuffle_seeds=np.array(seeds, dtype=np.uint64),
                            meta=np.array([root, k, min(shuffles), len(list(shuffles))],
                                          dtype=np.int64))
        print(name, len(rows), "games")

    # --- the reference's 12-game raw oracle (target 100, one max_rounds=0 override)
    profile = GameProfile(default_target_score=100, default_max_rounds=200,
                          tournament_max_rounds_overrides=(
                              TournamentMaxRoundsOverride(root_seed=11, k=2, shuffle_index=0,
                                                          game_index=0, max_rounds=0),))
    o12 = {}
    for root in (11, 22):
        for k in (2, 4):
            strats, rows, tl, seeds = play_shuffles("tiny", root, k, [0, 1], profile=profile)
            o12[f"rows_{root}_{k}"] = rows_to_compact(rows, k)
            o12[f"tallies_{root}_{k}"] = tl
    o12["strategies"] = pack_strategies(grid("tiny"))
    np.savez_compressed(HERE / "oracle12.npz", **o12)

    # --- public helper KAT
    hs = [ThresholdStrategy(score_threshold=0, dice_threshold=6),
          ThresholdStrategy(score_threshold=500, dice_threshold=3),
          ThresholdStrategy(score_threshold=1000, dice_threshold=2)]
    df = ref_sim.simulate_many_games(n_games=10, strategies=hs, target_score=5000, seed=123)
    assert df["winner_seat"].value_counts().to_dict() == {"P2": 6, "P1": 2, "P3": 2}
    np.savez_compressed(HERE / "helpers.npz", strategies=pack_strategies(hs),
                        rows=rows_to_compact(df.to_dict("records"), 3),
                        game_seeds=df["game_seed"].to_numpy(dtype=np.uint64))

    # --- H2H blocks
    from farkle.analysis.h2h_schedule import _simulate_block_from_manifest
    full = grid("full")
    manifest = build_strategy_manifest(full)
    never = [i for i, s in enumerate(full)
             if s.require_both and s.dice_threshold == 0 and s.consider_dice][:2]
    blocks = []
    for pair_id, (a, b), order, target, max_att in (
        (0, (0, 1), 0, 40, 80), (1, (17, 4000), 1, 40, 80), (2, (never[0], never[1]), 0, 5, 10),
        (3, (never[0], 2500), 1, 30, 60), (4, (5159, 123), 0, 25, 50),
    ):
        s1, s2 = (a, b) if order == 0 else (b, a)
        block = {"block_id": f"b{pair_id}", "root_seed": 42, "pair_id": pair_id, "order": order,
                 "seat1_strategy": s1, "seat2_strategy": s2, "n_completed_required": target,
                 "max_attempts": max_att, "rng_scheme_version": 2,
                 "rng_purpose_namespace": 203}
        # two chunks to exercise resume-from-progress
        p1 = _simulate_block_from_manifest(dict(block), manifest, 13)
        p2 = _simulate_block_from_manifest({**block, **{k_: p1[k_] for k_ in (
            "games_attempted", "games_completed", "games_safety_limit", "wins_seat1",
            "wins_seat2")}}, manifest, 5000)
        keep = ("games_attempted", "games_completed", "games_safety_limit", "wins_seat1",
                "wins_seat2")
        blocks.append({**block, "seat1": list(pack_strategy(full[s1])),
                       "seat2": list(pack_strategy(full[s2])),
                       "after_chunk13": {k_: int(p1[k_]) for k_ in keep},
                       "final": {k_: int(p2[k_]) for k_ in keep},
                       "completion_status": p2["completion_status"]})
    (HERE / "h2h.json").write_text(json.dumps(blocks, indent=0))

    # --- full fast grid, seed 42, k=2 (24,000 games)
    chunks = [(i, min(i + 25, 600)) for i in range(0, 600, 25)]
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        parts = list(ex.map(_fast42_chunk, chunks))
    tl = np.sum(parts, axis=0)
    np.savez_compressed(HERE / "fast42.npz", tallies=tl,
                        strategies=pack_strategies(grid("fast")))
    print("fast42 wins", tl[[42, 46, 51, 37, 25], 0], "attempted", tl[:, 1].sum() // 2,
          "completed", tl[:, 2].sum() // 2, "safety", tl[:, 3].sum() // 2)


if __name__ == "__main__":
    main()
