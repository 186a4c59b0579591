"""Game-level host surface: rows, ``_play_game`` and the ``simulate_many_games`` helpers.

Mirror of the reference's ``farkle.simulation.simulation`` for the hot path:

* ``PlayerRngCoordinates``               simulation.py:332-358
* ``_play_game``                         simulation.py:576-655
* ``simulate_many_games[_from_seeds]``   simulation.py:658-790
* ``simulation_rows_to_table``           simulation.py:565-573 with the Arrow schema of
                                         utils/schema_helpers.py:23-90

The games themselves are played by the CUDA library (``fb_play_games`` /
``fb_play_tournament``); this module only turns compact ``fb_row_*`` records into the
reference's flat row mapping (same keys, same key order, same Python types, ``None`` where
the reference writes nulls) or — for ingest-rate output — straight into an Arrow table.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Iterable, Mapping, Sequence

import numpy as np

from .layout import ROW_I16_OVERFLOW, ROW_ROLL_LIMIT, ROW_SAFETY_LIMIT, STRATEGY_DTYPE
from .random import RNG_SCHEME_VERSION, RandomPurpose, coordinate_entropy, spawn_seeds
from .strategies import ThresholdStrategy, pack_strategies, prepare_strategy_ids

OUTCOME_SCHEMA_VERSION = 2      # utils/schema_helpers.py:17
TOURNAMENT_METHOD_VERSION = 2   # utils/schema_helpers.py:18
DEFAULT_TARGET_SCORE = 10_000
DEFAULT_MAX_ROUNDS = 200

# PlayerStats field order (game/engine.py:385-398) = order of the per-seat row keys
SEAT_FIELDS = ("score", "farkles", "rolls", "n_turns", "highest_turn", "strategy", "rank",
               "loss_margin", "smart_five_uses", "n_smart_five_dice", "smart_one_uses",
               "n_smart_one_dice", "hot_dice", "hit_max_rounds")


class RollLimitError(RuntimeError):
    """A turn exceeded ROLL_LIMIT=1000 rolls (the reference raises at engine.py:242-243)."""


@dataclass(frozen=True, slots=True)
class PlayerRngCoordinates:
    """Complete semantic coordinates of a table's seat streams (simulation.py:332-358)."""

    purpose: RandomPurpose
    root_seed: int
    k: int
    shuffle_index: int = 0
    pair_id: int = 0
    order: int = 0
    game_index: int | None = None
    attempt_index: int | None = None

    def coords7(self) -> list[int]:
        """``fb_play_games`` coordinate row {purpose, root, k, shuffle, pair, order, game}."""
        e = coordinate_entropy(self.purpose, root_seed=self.root_seed, k=self.k,
                               shuffle_index=self.shuffle_index, pair_id=self.pair_id,
                               order=self.order, game_index=self.game_index,
                               attempt_index=self.attempt_index)
        v = [e[2 + 2 * i] | (e[3 + 2 * i] << 32) for i in range(6)]
        return [int(self.purpose), *v]


def _prepare_public_helper_strategies(strategies: Sequence[ThresholdStrategy]
                                      ) -> list[ThresholdStrategy]:
    """Copies with one unique canonical id per seat (simulation.py:361-409)."""
    from dataclasses import replace

    ids = prepare_strategy_ids(strategies)
    return [replace(s, strategy_id=i) for s, i in zip(strategies, ids, strict=True)]


# --------------------------------------------------------------------------- row expansion
def check_row_flags(rows: np.ndarray) -> None:
    """Surface the conditions under which the reference raises instead of returning a row."""
    flags = rows["flags"]
    if (flags & ROW_ROLL_LIMIT).any():
        raise RollLimitError("Roll limit reached in a turn")  # engine.py:242-243
    if (flags & ROW_I16_OVERFLOW).any():
        raise OverflowError("a per-seat counter left the int16 range of the row schema "
                            "(utils/schema_helpers.py:44-58); the reference's Arrow conversion raises")


def validate_compact_rows(rows: np.ndarray) -> None:
    """The closed outcome invariants of ``validate_simulation_row`` (simulation.py:450-563) that a
    compact row can violate, checked for a whole launch at once: distinct seated strategies, a
    completed game's winner is its rank-1 seat (highest score, ties to the lower seat), a
    safety-limit game claims no winner.  Every other field the reference validates (ranks, margins,
    seat order, null patterns) is DERIVED from these columns by ``expand_rows`` /
    ``compact_rows_to_table``, so it cannot disagree with them."""
    n = len(rows)
    if n == 0:
        return
    seats = rows["seats"]
    k = seats.shape[1]
    strategies = np.sort(seats["strategy"].astype(np.int64), axis=1)
    if k > 1 and (strategies[:, 1:] == strategies[:, :-1]).any():
        raise ValueError("Simulation row must seat distinct strategies")
    safety = (rows["flags"] & ROW_SAFETY_LIMIT) != 0
    scores = seats["score"].astype(np.int64)
    best = np.argmax(scores, axis=1)                      # first maximum = lowest seat among ties
    winner = rows["winner_seat"].astype(np.int64)
    if (~safety & (winner != best)).any():
        raise ValueError("Completed simulation row must have exactly one winner matching its rank-1 seat")
    if (safety & (winner != 0xFF)).any():
        raise ValueError("Safety-limit simulation row cannot claim a winner: ['winner_seat']")


def validate_simulation_row(row: Mapping[str, Any]) -> None:
    """Validate one flattened game row: the reference's closed outcome invariants, check for check
    and message for message (simulation/simulation.py:450-563)."""
    try:
        n_players = int(row["k"])
        status = str(getattr(row["termination_status"], "value", row["termination_status"]))
        if status not in ("completed", "safety_limit"):
            raise ValueError(status)
    except (KeyError, TypeError, ValueError) as exc:
        raise ValueError("Simulation row has invalid k or termination_status") from exc
    if n_players < 1:
        raise ValueError("Simulation row k must be positive")
    if row.get("outcome_schema_version") != OUTCOME_SCHEMA_VERSION:
        raise ValueError(f"Simulation row must use outcome_schema_version={OUTCOME_SCHEMA_VERSION}")
    seats = [f"P{index}" for index in range(1, n_players + 1)]
    strategies: list[int] = []
    scores: list[int] = []
    for seat in seats:
        if row.get(f"{seat}_strategy") is None:
            raise ValueError(f"Simulation row missing seated strategy {seat}_strategy")
        sid = row[f"{seat}_strategy"]
        if isinstance(sid, bool) or not isinstance(sid, (int, np.integer)):
            raise ValueError(f"simulation row {seat}_strategy must be a canonical integer strategy id")
        strategies.append(int(sid))
        score = row.get(f"{seat}_score")
        if isinstance(score, bool) or not isinstance(score, (int, np.integer)):
            raise ValueError(f"Simulation row {seat}_score must be an integer")
        scores.append(int(score))
    if len(set(strategies)) != n_players:
        raise ValueError("Simulation row must seat distinct strategies")
    missing_ranks = [seat for seat in seats if f"{seat}_rank" not in row]
    if missing_ranks:
        raise ValueError(f"Simulation row missing participant ranks for {missing_ranks}")
    ranks = [row[f"{seat}_rank"] for seat in seats]
    winner_seat, winner_strategy = row.get("winner_seat"), row.get("winner_strategy")
    if status == "completed":
        rank_one = [seat for seat, rank in zip(seats, ranks, strict=True) if rank == 1]
        if not isinstance(winner_seat, str) or len(rank_one) != 1 or winner_seat != rank_one[0]:
            raise ValueError("Completed simulation row must have exactly one winner matching its rank-1 seat")
        if any(rank is None for rank in ranks) or sorted(ranks) != list(range(1, n_players + 1)):
            raise ValueError("Completed simulation row ranks must be the permutation 1..k")
        score_order = sorted(range(n_players), key=lambda index: (-scores[index], index))
        expected = [0] * n_players
        for rank, index in enumerate(score_order, start=1):
            expected[index] = rank
        if [int(rank) for rank in ranks] != expected:
            raise ValueError("Completed simulation row ranks are inconsistent with final scores")
        if winner_strategy is None or winner_strategy != row.get(f"{winner_seat}_strategy"):
            raise ValueError("Completed simulation row must identify the winning strategy")
        winning_score, victory_margin = row.get("winning_score"), row.get("victory_margin")
        if winning_score is None or victory_margin is None:
            raise ValueError("Completed simulation row must retain winner-conditioned fields")
        if row.get("hit_safety_limit") is not False:
            raise ValueError("Completed simulation row cannot hit the safety limit")
        if any(row.get(f"{seat}_hit_max_rounds") is not False for seat in seats):
            raise ValueError("Completed simulation row cannot mark a seat at the safety limit")
        if int(winning_score) != scores[seats.index(winner_seat)] or int(winning_score) != max(scores):
            raise ValueError("Completed simulation row has inconsistent winning_score")
        ordered = sorted(scores, reverse=True)
        if int(victory_margin) != int(winning_score) - (ordered[1] if n_players > 1 else 0):
            raise ValueError("Completed simulation row has inconsistent victory_margin")
        for seat, score in zip(seats, scores, strict=True):
            margin = row.get(f"{seat}_loss_margin")
            if (isinstance(margin, bool) or not isinstance(margin, (int, np.integer))
                    or int(margin) != int(winning_score) - score):
                raise ValueError(f"Completed simulation row has inconsistent {seat}_loss_margin")
        seat_ranks = row.get("seat_ranks")
        if seat_ranks is None or list(seat_ranks) != [seats[index] for index in score_order]:
            raise ValueError("Completed simulation row has inconsistent seat_ranks")
        return
    if row.get("hit_safety_limit") is not True:
        raise ValueError("Safety-limit simulation row must set hit_safety_limit=true")
    if any(row.get(f"{seat}_hit_max_rounds") is not True for seat in seats):
        raise ValueError("Safety-limit simulation row must mark every seat at the safety limit")
    present = [name for name, value in (("winner_seat", winner_seat), ("winner_strategy", winner_strategy),
                                        ("winning_score", row.get("winning_score")),
                                        ("victory_margin", row.get("victory_margin"))) if value is not None]
    if present:
        raise ValueError(f"Safety-limit simulation row cannot claim a winner: {present}")
    if any(rank is not None for rank in ranks):
        raise ValueError("Safety-limit simulation row cannot assign participant ranks")
    seat_ranks = row.get("seat_ranks")
    if seat_ranks is None or list(seat_ranks) != [None] * n_players:
        raise ValueError("Safety-limit simulation row must retain k null seat-rank entries")
    if any(row.get(f"{seat}_loss_margin") is not None for seat in seats):
        raise ValueError("Safety-limit simulation row cannot assign loss margins")


def derive_ranks(rows: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """``(rank[n, k] (1-based, 0 = null), order[n, k])`` — stable sort by score desc, seat asc
    (game/engine.py:483), nulls for safety-limit rows."""
    scores = rows["seats"]["score"].astype(np.int64)
    n, k = scores.shape
    # rank of seat s = 1 + seats ahead of it (higher score, or equal score and lower seat): k*k
    # vector comparisons instead of n small sorts; order = the inverse permutation of rank
    rank = np.ones((n, k), dtype=np.int8)
    for s in range(k):
        for o in range(k):
            if o != s:
                ahead = scores[:, o] > scores[:, s] if o > s else scores[:, o] >= scores[:, s]
                rank[:, s] += ahead
    order = np.empty((n, k), dtype=np.intp)
    np.put_along_axis(order, rank.astype(np.intp) - 1, np.arange(k, dtype=np.intp)[None, :], axis=1)
    safety = (rows["flags"] & ROW_SAFETY_LIMIT) != 0
    rank[safety] = 0
    return rank, order


def expand_rows(rows: np.ndarray, provenance: Sequence[Mapping[str, Any]] | None = None,
                *, root_seed: Sequence[int] | int = 0,
                purpose_namespace: int = int(RandomPurpose.INDEXED_SEED)) -> list[dict[str, Any]]:
    """Compact rows -> the reference's flat row mappings (simulation.py:612-654).

    ``provenance[i]`` is applied exactly like ``flat.update(provenance)``; without it the
    defaults of ``_play_game`` are used (``root_seed``/``game_seed`` = the seed argument).
    """
    check_row_flags(rows)
    n = len(rows)
    k = rows["seats"].shape[1] if n else 0
    rank, order = derive_ranks(rows) if n else (None, None)
    seeds = np.broadcast_to(np.asarray(root_seed, dtype=np.uint64), (n,))
    out: list[dict[str, Any]] = []
    for i in range(n):
        r = rows[i]
        safety = bool(r["flags"] & ROW_SAFETY_LIMIT)
        seats = r["seats"]
        scores = [int(x) for x in seats["score"]]
        if safety:
            winner, winner_strategy, seat_ranks = None, None, [None] * k
            winning_score = margin = None
        else:
            w = int(r["winner_seat"])
            winner = f"P{w + 1}"
            winner_strategy = int(seats["strategy"][w])
            seat_ranks = [f"P{int(s) + 1}" for s in order[i]]
            winning_score = scores[w]
            ordered = sorted(scores, reverse=True)
            margin = ordered[0] - (ordered[1] if k > 1 else 0)
        flat: dict[str, Any] = {
            "termination_status": "safety_limit" if safety else "completed",
            "hit_safety_limit": safety,
            "outcome_schema_version": OUTCOME_SCHEMA_VERSION,
            "winner_seat": winner,
            "winner_strategy": winner_strategy,
            "seat_ranks": seat_ranks,
            "winning_score": winning_score,
            "victory_margin": margin,
            "n_rounds": int(r["n_rounds"]),
            "root_seed": int(seeds[i]),
            "k": k,
            "shuffle_index": None,
            "game_index": None,
            "deterministic_batch_id": None,
            "game_seed": int(seeds[i]),
            "rng_scheme_version": RNG_SCHEME_VERSION,
            "rng_purpose_namespace": purpose_namespace,
        }
        if provenance is not None:
            flat.update(provenance[i])
        # per-seat blocks follow the iteration order of GameMetrics.players: rank order for a
        # completed game, seat order at the safety limit (game/engine.py:477-509)
        for s in (range(k) if safety else (int(x) for x in order[i])):
            seat = seats[s]
            p = f"P{s + 1}_"
            flat[p + "score"] = scores[s]
            flat[p + "farkles"] = int(seat["farkles"])
            flat[p + "rolls"] = int(seat["rolls"])
            flat[p + "n_turns"] = int(seat["n_turns"])
            flat[p + "highest_turn"] = int(seat["highest_turn"])
            flat[p + "strategy"] = int(seat["strategy"])
            flat[p + "rank"] = None if safety else int(rank[i, s])
            flat[p + "loss_margin"] = None if safety else winning_score - scores[s]
            flat[p + "smart_five_uses"] = int(seat["smart_five_uses"])
            flat[p + "n_smart_five_dice"] = int(seat["n_smart_five_dice"])
            flat[p + "smart_one_uses"] = int(seat["smart_one_uses"])
            flat[p + "n_smart_one_dice"] = int(seat["n_smart_one_dice"])
            flat[p + "hot_dice"] = int(seat["hot_dice"])
            flat[p + "hit_max_rounds"] = safety
        validate_simulation_row(flat)          # as `_play_game` does before returning a row (:655)
        out.append(flat)
    return out


def raw_simulation_schema_for(n_players: int):
    """Arrow schema of persisted rows (utils/schema_helpers.py:23-90), field for field."""
    import pyarrow as pa

    if n_players < 1:
        raise ValueError("n_players must be positive")
    str_list = pa.list_(pa.field("item", pa.string(), nullable=True))
    base = [
        pa.field("root_seed", pa.int64(), nullable=False),
        pa.field("k", pa.int16(), nullable=False),
        pa.field("shuffle_index", pa.int64(), nullable=False),
        pa.field("game_index", pa.int32(), nullable=False),
        pa.field("deterministic_batch_id", pa.int32(), nullable=False),
        pa.field("shuffle_seed", pa.int64(), nullable=False),
        pa.field("termination_status", pa.string(), nullable=False),
        pa.field("hit_safety_limit", pa.bool_(), nullable=False),
        pa.field("outcome_schema_version", pa.int16(), nullable=False),
        pa.field("winner_seat", pa.string(), nullable=True),
        pa.field("winner_strategy", pa.int32(), nullable=True),
        pa.field("game_seed", pa.int64(), nullable=False),
        pa.field("rng_scheme_version", pa.int16(), nullable=False),
        pa.field("rng_purpose_namespace", pa.int32(), nullable=False),
        pa.field("seat_ranks", str_list, nullable=False),
        pa.field("winning_score", pa.int32(), nullable=True),
        pa.field("victory_margin", pa.int32(), nullable=True),
        pa.field("n_rounds", pa.int16(), nullable=False),
    ]
    seat = {
        "score": (pa.int32(), False), "farkles": (pa.int16(), False), "rolls": (pa.int16(), False),
        "highest_turn": (pa.int16(), False), "strategy": (pa.int32(), False),
        "rank": (pa.int8(), True), "loss_margin": (pa.int32(), True),
        "smart_five_uses": (pa.int16(), False), "n_smart_five_dice": (pa.int16(), False),
        "smart_one_uses": (pa.int16(), False), "n_smart_one_dice": (pa.int16(), False),
        "hot_dice": (pa.int16(), False), "n_turns": (pa.int16(), False),
        "hit_max_rounds": (pa.bool_(), False),
    }
    fields = [pa.field(f"P{i}_{name}", t, nullable=nullable)
              for i in range(1, n_players + 1) for name, (t, nullable) in seat.items()]
    return pa.schema([*base, *fields])


def simulation_rows_to_table(rows: Sequence[Mapping[str, Any]], n_players: int):
    """Row mappings -> Arrow table (simulation.py:565-573)."""
    import pyarrow as pa

    for row in rows:
        if int(row["k"]) != n_players:
            raise ValueError(f"Simulation row k={row['k']} does not match schema k={n_players}")
    return pa.Table.from_pylist(list(rows), schema=raw_simulation_schema_for(n_players))


def compact_rows_to_table(rows: np.ndarray, *, root_seed: int, k: int, shuffle_index,
                          game_index, deterministic_batch_id, shuffle_seed,
                          purpose_namespace: int = int(RandomPurpose.TOURNAMENT_GAME)):
    """Vectorised compact rows -> Arrow table with the schema above (no Python per-row loop).

    Equal, column for column, to ``simulation_rows_to_table(expand_rows(...))``; this is the
    ingest-rate path for tournament row shards.  ``shuffle_index`` / ``game_index`` /
    ``deterministic_batch_id`` / ``shuffle_seed`` are scalars or per-row arrays.
    """
    import pyarrow as pa

    check_row_flags(rows)
    validate_compact_rows(rows)
    n = len(rows)
    schema = raw_simulation_schema_for(k)
    safety = (rows["flags"] & ROW_SAFETY_LIMIT) != 0
    rank, order = derive_ranks(rows)
    seats = rows["seats"]
    scores = seats["score"].astype(np.int32)
    w = np.where(safety, 0, rows["winner_seat"]).astype(np.int64)
    ar = np.arange(n)
    win_score = scores[ar, w]
    margin = scores[ar, order[:, 0]] - (scores[ar, order[:, 1]] if k > 1 else 0)   # winner minus runner-up
    # string columns by `take` from tiny dictionaries (no Python string objects per row: at ingest
    # rate this build, not the GPU, is the bottleneck of rows mode)
    names = pa.array([f"P{i + 1}" for i in range(k)], type=pa.string())
    status_names = pa.array(["completed", "safety_limit"], type=pa.string())

    def full(v, dtype):
        return np.broadcast_to(np.asarray(v, dtype=dtype), (n,))

    seat_rank_values = names.take(pa.array(order.reshape(-1).astype(np.int32), mask=np.repeat(safety, k)))
    seat_ranks = pa.ListArray.from_arrays(pa.array(np.arange(0, n * k + 1, k, dtype=np.int32)),
                                          seat_rank_values, type=schema.field("seat_ranks").type)
    cols: dict[str, Any] = {
        "root_seed": pa.array(full(root_seed, np.int64)),
        "k": pa.array(full(k, np.int16)),
        "shuffle_index": pa.array(full(shuffle_index, np.int64)),
        "game_index": pa.array(full(game_index, np.int32)),
        "deterministic_batch_id": pa.array(full(deterministic_batch_id, np.int32)),
        "shuffle_seed": pa.array(full(shuffle_seed, np.int64)),
        "termination_status": status_names.take(pa.array(safety.astype(np.int8))),
        "hit_safety_limit": pa.array(safety),
        "outcome_schema_version": pa.array(full(OUTCOME_SCHEMA_VERSION, np.int16)),
        "winner_seat": names.take(pa.array(w.astype(np.int32), mask=safety)),
        "winner_strategy": pa.array(seats["strategy"][ar, w].astype(np.int32), mask=safety),
        "game_seed": pa.array(rows["game_seed"].astype(np.int64)),
        "rng_scheme_version": pa.array(full(RNG_SCHEME_VERSION, np.int16)),
        "rng_purpose_namespace": pa.array(full(purpose_namespace, np.int32)),
        "seat_ranks": seat_ranks,
        "winning_score": pa.array(win_score, mask=safety),
        "victory_margin": pa.array(margin.astype(np.int32), mask=safety),
        "n_rounds": pa.array(rows["n_rounds"].astype(np.int16)),
    }
    for s in range(k):
        p = f"P{s + 1}_"
        seat = seats[:, s]
        cols[p + "score"] = pa.array(scores[:, s])
        cols[p + "farkles"] = pa.array(seat["farkles"].astype(np.int16))
        cols[p + "rolls"] = pa.array(seat["rolls"].astype(np.int16))
        cols[p + "highest_turn"] = pa.array(seat["highest_turn"].astype(np.int16))
        cols[p + "strategy"] = pa.array(seat["strategy"].astype(np.int32))
        cols[p + "rank"] = pa.array(rank[:, s], mask=safety)
        cols[p + "loss_margin"] = pa.array((win_score - scores[:, s]).astype(np.int32), mask=safety)
        cols[p + "smart_five_uses"] = pa.array(seat["smart_five_uses"].astype(np.int16))
        cols[p + "n_smart_five_dice"] = pa.array(seat["n_smart_five_dice"].astype(np.int16))
        cols[p + "smart_one_uses"] = pa.array(seat["smart_one_uses"].astype(np.int16))
        cols[p + "n_smart_one_dice"] = pa.array(seat["n_smart_one_dice"].astype(np.int16))
        cols[p + "hot_dice"] = pa.array(seat["hot_dice"].astype(np.int16))
        cols[p + "n_turns"] = pa.array(seat["n_turns"].astype(np.int16))
        cols[p + "hit_max_rounds"] = pa.array(safety)
    return pa.Table.from_arrays([cols[f.name] for f in schema], schema=schema)


# --------------------------------------------------------------------------- games
def play_games_batch(strategy_rows: Sequence[Sequence[ThresholdStrategy]],
                     coordinates: Sequence[PlayerRngCoordinates], *,
                     target_score: int | Sequence[int] = DEFAULT_TARGET_SCORE,
                     max_rounds: int | Sequence[int] = DEFAULT_MAX_ROUNDS,
                     device: int | None = None) -> np.ndarray:
    """Play ``len(coordinates)`` independent tables in ONE launch; returns compact rows."""
    from .device import get_engine

    n = len(coordinates)
    if n == 0:
        raise ValueError("no games requested")
    k = len(strategy_rows[0])
    table = np.empty((n, k), dtype=STRATEGY_DTYPE)
    ids = np.empty((n, k), dtype=np.int32)
    for i, (strats, c) in enumerate(zip(strategy_rows, coordinates, strict=True)):
        if len(strats) != k or c.k != k:
            raise ValueError(
                "Player RNG coordinate k does not match the number of seated strategies")
        table[i] = pack_strategies(strats)
        sid = [s.strategy_id for s in strats]
        if any(x is None for x in sid):
            raise ValueError("every seated strategy needs a strategy_id "
                             "(use _prepare_public_helper_strategies)")
        if len(set(sid)) != k:
            raise ValueError("Simulation row must seat distinct strategies")
        ids[i] = sid
    coords = np.array([c.coords7() for c in coordinates], dtype=np.uint64)
    scalar_t, scalar_m = np.isscalar(target_score), np.isscalar(max_rounds)
    rows, _totals = get_engine(device).play_games(
        coords, k, table, seat_strategy_ids=ids,
        target_score=int(target_score) if scalar_t else DEFAULT_TARGET_SCORE,
        max_rounds=int(max_rounds) if scalar_m else DEFAULT_MAX_ROUNDS,
        target_scores=None if scalar_t else np.asarray(target_score, dtype=np.int32),
        max_rounds_v=None if scalar_m else np.asarray(max_rounds, dtype=np.int32))
    return rows


def _play_game(seed: int, strategies: Sequence[ThresholdStrategy], target_score: int = 10_000,
               provenance: Mapping[str, Any] | None = None, max_rounds: int = 200,
               player_rng_coordinates: PlayerRngCoordinates | None = None,
               *, device: int | None = None) -> Mapping[str, Any]:
    """Play a single game on the GPU and return the reference's flat row (simulation.py:576-655)."""
    coords = player_rng_coordinates or PlayerRngCoordinates(
        purpose=RandomPurpose.PLAYER, root_seed=seed, k=len(strategies))
    if coords.k != len(strategies):
        raise ValueError("Player RNG coordinate k does not match the number of seated strategies")
    rows = play_games_batch([strategies], [coords], target_score=target_score,
                            max_rounds=max_rounds, device=device)
    return expand_rows(rows, None if provenance is None else [provenance], root_seed=seed)[0]


def _helper_rows(seeds: Iterable[int], strategies: Sequence[ThresholdStrategy], target_score: int,
                 root_seed: int | None, device: int | None) -> list[dict[str, Any]]:
    resolved = _prepare_public_helper_strategies(strategies)
    k = len(resolved)
    seeds = [int(s) for s in seeds]
    if not seeds:
        return []
    coords, prov = [], []
    for game_index, game_seed in enumerate(seeds):
        root = game_seed if root_seed is None else root_seed
        coords.append(PlayerRngCoordinates(
            purpose=RandomPurpose.PLAYER, root_seed=root, k=k,
            game_index=0 if root_seed is None else game_index))
        prov.append({"root_seed": root, "k": k, "shuffle_index": None, "game_index": game_index,
                     "deterministic_batch_id": None, "game_seed": game_seed,
                     "rng_scheme_version": RNG_SCHEME_VERSION,
                     "rng_purpose_namespace": int(RandomPurpose.INDEXED_SEED)})
    rows = play_games_batch([resolved] * len(seeds), coords, target_score=target_score,
                            device=device)
    return expand_rows(rows, prov, root_seed=seeds)


def simulate_many_games(*, n_games: int, strategies: Sequence[ThresholdStrategy],
                        target_score: int = 10_000, seed: int | None = None, n_jobs: int = 1,
                        device: int | None = None):
    """``n_games`` games of one table as a DataFrame (simulation.py:658-722).

    ``n_jobs`` is accepted for signature compatibility; all games run in one launch.
    """
    import pandas as pd

    if seed is None:
        raise ValueError("simulate_many_games requires an explicit seed")
    del n_jobs
    return pd.DataFrame(_helper_rows(spawn_seeds(n_games, seed=seed), strategies, target_score,
                                     seed, device))


def simulate_many_games_from_seeds(*, seeds: Iterable[int], strategies: Sequence[ThresholdStrategy],
                                   target_score: int = 10_000, n_jobs: int = 1,
                                   root_seed: int | None = None, device: int | None = None):
    """Games for predetermined seeds (simulation.py:725-790)."""
    import pandas as pd

    del n_jobs
    return pd.DataFrame(_helper_rows(seeds, strategies, target_score, root_seed, device))


def aggregate_metrics(df) -> Mapping[str, Any]:
    """Summary of a results frame (simulation.py:823-838)."""
    return {"games": len(df), "avg_rounds": df["n_rounds"].mean(),
            "winner_freq": df["winner_seat"].value_counts().to_dict()}
