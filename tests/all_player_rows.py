"""Test infrastructure: the unconditional all-player statistics the slow way, from compact rows.

A plain numpy restatement of analysis/all_player_metrics.py:262-340 (``_seat_exposure_columns`` /
``_update_exposure_columns``): rows in source order, seats within a row in seat order, float64
accumulation with unbuffered ``np.add.at``.  The device kernel (``allplayer_gather_kernel``) is
checked against it, and it is itself checked against the reference's own stage function on a
Parquet of the same rows (tests/test_reference_dropin.py).
"""

from __future__ import annotations

import numpy as np

from farkle_ii_b200.layout import ALLP_WIDTH


def all_player_from_rows(rows: np.ndarray, slot_of_row: np.ndarray, n_slots: int, n_ids: int) -> np.ndarray:
    """``int64 [n_slots, n_ids, ALLP_WIDTH]`` (columns 41..44 hold float64 bit patterns)."""
    k = rows["seats"].shape[1]
    safety = (rows["flags"] & 1) != 0
    completed = ~safety
    rounds = rows["n_rounds"].astype(np.float64)
    seats = rows["seats"]
    scores_all = seats["score"].astype(np.int64)
    out_f = np.zeros((n_slots, n_ids, ALLP_WIDTH), dtype=np.float64)
    cols: list[np.ndarray] = []
    strategies = []
    for s in range(k):
        seat = seats[:, s]
        score = seat["score"].astype(np.float64)
        turns = seat["n_turns"].astype(np.float64)
        won = completed & (rows["winner_seat"] == s)
        exact = np.divide(score, turns, out=np.zeros_like(score), where=turns != 0)
        proxy = np.divide(score, rounds, out=np.zeros_like(score), where=rounds != 0)
        diff = turns - rounds
        mine = scores_all[:, s]
        ahead = np.zeros(len(rows), dtype=np.int64)
        for o in range(k):
            ahead += (scores_all[:, o] > mine) | ((scores_all[:, o] == mine) & (o < s))
        rank = (ahead + 1).astype(np.float64)
        margin = (scores_all.max(axis=1) - mine).astype(np.float64)
        behaviours = [rank, margin, seat["rolls"], seat["farkles"], seat["highest_turn"], seat["hot_dice"],
                      seat["smart_five_uses"], seat["n_smart_five_dice"], seat["smart_one_uses"],
                      seat["n_smart_one_dice"]]
        c = np.zeros((len(rows), ALLP_WIDTH), dtype=np.float64)
        c[:, 0] = 1
        c[:, 1] = completed
        c[:, 2] = safety
        c[:, 3] = won
        c[:, 4] = diff != 0
        c[:, 5], c[:, 6] = score, score * score
        c[:, 7], c[:, 8] = turns, turns * turns
        c[:, 9], c[:, 10] = diff, diff * diff
        for b, values in enumerate(behaviours):
            present = completed if b < 2 else np.ones(len(rows), dtype=bool)
            v = np.where(present, np.asarray(values, dtype=np.float64), 0.0)
            c[:, 11 + 3 * b], c[:, 12 + 3 * b], c[:, 13 + 3 * b] = present, v, v * v
        c[:, 41], c[:, 42], c[:, 43], c[:, 44] = exact, exact * exact, proxy, proxy * proxy
        cols.append(c)
        strategies.append(seat["strategy"].astype(np.int64))
    values = np.stack(cols, axis=1).reshape(-1, ALLP_WIDTH)          # row-major: row 0 seat 0, row 0 seat 1, ...
    sid = np.stack(strategies, axis=1).reshape(-1)
    slot = np.repeat(np.asarray(slot_of_row, dtype=np.int64), k)
    flat = out_f.reshape(n_slots * n_ids, ALLP_WIDTH)
    for col in range(ALLP_WIDTH):
        np.add.at(flat[:, col], slot * n_ids + sid, values[:, col])
    out = np.empty((n_slots, n_ids, ALLP_WIDTH), dtype=np.int64)
    out[..., :41] = out_f[..., :41].astype(np.int64)
    out[..., 41:] = out_f[..., 41:].view(np.int64)
    return out
