#!/usr/bin/env python
"""Golden fixtures for the host-surface mirror, generated from the REAL reference.

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden_host.py

Freezes, as JSON, what the reference's own seam functions return so the mirror in
farkle_ii_b200/{run_tournament,simulation,h2h}.py can be compared key for key:

* ``host_surface.json``
    - ``shuffles``:  _play_one_shuffle(task, collect_rows=True) -> wins / outcome payload /
                     sums / sq_sums / full row dicts (fast grid root 54 k=4 shuffles 3,4; the
                     tiny grid root 11 k=2 shuffles 0,1 with the max_rounds=0 override)
    - ``chunk``:     _run_chunk / _run_chunk_metrics over those shuffles
    - ``helpers``:   simulate_many_games / simulate_many_games_from_seeds / _play_game rows
    - ``h2h``:       complete _simulate_block_from_manifest result dicts (two chunks per block)
    - ``schema``:    raw_simulation_schema_for(k) field list for k = 2, 4
Only values the reference computes are stored; no reference source is copied.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.path.insert(0, "/root/reference/src")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import numpy as np  # noqa: E402

import make_golden as mg  # noqa: E402
from farkle.analysis.h2h_schedule import _simulate_block_from_manifest  # noqa: E402
from farkle.simulation import run_tournament as rt  # noqa: E402
from farkle.simulation import simulation as ref_sim  # noqa: E402
from farkle.simulation.game_profile import (  # noqa: E402
    GameProfile, H2HMaxRoundsOverride, TournamentMaxRoundsOverride)
from farkle.simulation.strategies import ThresholdStrategy, build_strategy_manifest  # noqa: E402
from farkle.utils import random as ur  # noqa: E402
from farkle.utils.schema_helpers import raw_simulation_schema_for  # noqa: E402


def js(o):
    if isinstance(o, dict):
        return {str(k): js(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [js(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    return o


def task(root, k, sh, bs):
    return rt.ShuffleTask(root, k, sh, ur.coordinate_seed(ur.RandomPurpose.TOURNAMENT_SHUFFLE,
                                                          root_seed=root, k=k, shuffle_index=sh,
                                                          dtype=np.uint32), sh // bs)


def counter_payload(w):
    return {"wins": dict(w), "outcome": w.outcome_payload()}


def main():
    out = {}
    profile = GameProfile(default_target_score=100, default_max_rounds=200,
                          tournament_max_rounds_overrides=(
                              TournamentMaxRoundsOverride(root_seed=11, k=2, shuffle_index=0,
                                                          game_index=0, max_rounds=0),))
    cases = []
    for kind, root, k, shuffles, prof in (("fast", 54, 4, [3, 4], None), ("tiny", 11, 2, [0, 1], profile)):
        strats = mg.grid(kind)
        cfg = rt.TournamentConfig(n_players=k, n_strategies=len(strats))
        rt._init_worker(strats, cfg, prof)
        tasks = [task(root, k, s, 2) for s in shuffles]
        per = []
        for t in tasks:
            w, s, q, rows = rt._play_one_shuffle(t, collect_rows=True)
            per.append({"task": [t.root_seed, t.k, t.shuffle_index, t.shuffle_seed, t.deterministic_batch_id],
                        **counter_payload(w), "sums": {m: dict(v) for m, v in s.items()},
                        "sq_sums": {m: dict(v) for m, v in q.items()}, "rows": rows})
        cw = rt._run_chunk(tasks)
        mw, ms, mq = rt._run_chunk_metrics(tasks)
        cases.append({"grid": kind, "root": root, "k": k, "profile": prof is not None,
                      "shuffles": per, "chunk": counter_payload(cw),
                      "chunk_metrics": {**counter_payload(mw), "sums": {m: dict(v) for m, v in ms.items()},
                                        "sq_sums": {m: dict(v) for m, v in mq.items()}}})
    out["cases"] = cases

    hs = [ThresholdStrategy(score_threshold=0, dice_threshold=6),
          ThresholdStrategy(score_threshold=500, dice_threshold=3),
          ThresholdStrategy(score_threshold=1000, dice_threshold=2, strategy_id=1)]
    df = ref_sim.simulate_many_games(n_games=6, strategies=hs, target_score=5000, seed=123)
    df2 = ref_sim.simulate_many_games_from_seeds(seeds=[5, 6, 7], strategies=hs[:2], target_score=3000)
    df3 = ref_sim.simulate_many_games_from_seeds(seeds=[5, 6, 7], strategies=hs[:2], target_score=3000,
                                                 root_seed=99)
    one = ref_sim._play_game(7, ref_sim._prepare_public_helper_strategies(hs), target_score=2000)
    lim = ref_sim._play_game(7, ref_sim._prepare_public_helper_strategies(hs[:2]), max_rounds=3)
    out["helpers"] = {"many": df.to_dict("records"), "from_seeds": df2.to_dict("records"),
                      "from_seeds_root": df3.to_dict("records"), "one": dict(one), "limited": dict(lim)}

    full = mg.grid("full")
    manifest = build_strategy_manifest(full)
    never = [i for i, s in enumerate(full) if s.require_both and s.dice_threshold == 0 and s.consider_dice][:2]
    h2h_profile = GameProfile(h2h_max_rounds_overrides=(
        H2HMaxRoundsOverride(root_seed=42, pair_id=4, order=0, attempt_index=2, max_rounds=1),))
    h2h = []
    for pair_id, (a, b), order, target, max_att, prof in (
            (0, (0, 1), 0, 40, 80, None), (1, (17, 4000), 1, 40, 80, None),
            (2, (never[0], never[1]), 0, 5, 10, None), (3, (never[0], 2500), 1, 30, 60, None),
            (4, (5159, 123), 0, 25, 50, h2h_profile)):
        s1, s2 = (a, b) if order == 0 else (b, a)
        block = {"block_id": f"b{pair_id}", "family_hash": "f" * 8, "schedule_hash": "s" * 8,
                 "root_seed": 42, "pair_id": pair_id, "order": order, "seat1_strategy": s1,
                 "seat2_strategy": s2, "n_completed_required": target, "max_attempts": max_att,
                 "rng_scheme_version": 2, "rng_purpose_namespace": 203, "_private": 1}
        p1 = _simulate_block_from_manifest(dict(block), manifest, 13, prof)
        p2 = _simulate_block_from_manifest(dict(p1), manifest, 5000, prof)
        h2h.append({"block": block, "profile": prof is not None, "after13": p1, "final": p2,
                    "seat1": list(mg.pack_strategy(full[s1])), "seat2": list(mg.pack_strategy(full[s2]))})
    out["h2h"] = h2h
    out["schema"] = {str(k): [[f.name, str(f.type), f.nullable] for f in raw_simulation_schema_for(k)]
                     for k in (2, 4)}
    (HERE / "host_surface.json").write_text(json.dumps(js(out), separators=(",", ":")))
    print("wrote", (HERE / "host_surface.json").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
