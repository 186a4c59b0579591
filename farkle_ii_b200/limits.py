"""Per-game limits: the two launch scalars plus a keyed table of ``max_rounds`` exceptions.

Plays the part of the reference's ``farkle.simulation.game_profile`` (src/farkle/simulation/
game_profile.py:24-200) for this engine: the same public classes, the same rejections and
the same identity hash, but organised around what a launch needs -- ``target_score`` and
``max_rounds`` go to the kernel as scalars, and the exceptions inside one launch's shuffle range
become the ``override_*`` arrays of ``fb_play_tournament``.  Lookups go through dict indexes built
once per profile instead of a scan per game.
"""

from __future__ import annotations

import hashlib
import json
from dataclasses import asdict, dataclass, field, fields
from typing import ClassVar

GAME_PROFILE_CONTRACT_VERSION = 1

Coordinate = tuple[int, int, int, int]


def _is_count(value: object) -> bool:
    """A plain non-negative int; bool is refused although it subclasses int."""
    return type(value) is not bool and isinstance(value, int) and value >= 0


def _check_count(value: object, what: str) -> None:
    if not _is_count(value):
        raise ValueError(f"{what} must be a non-negative integer")


class _Exception4:
    """Shared behaviour of the two exception records: four key fields, then ``max_rounds``."""

    __slots__ = ()
    _key: ClassVar[tuple[str, str, str, str]]

    def _validate(self) -> None:
        for name in self._key:
            _check_count(getattr(self, name), name)
        _check_count(self.max_rounds, "max_rounds")  # type: ignore[attr-defined]

    @property
    def coordinate(self) -> Coordinate:
        a, b, c, d = (getattr(self, name) for name in self._key)
        return (a, b, c, d)


@dataclass(frozen=True, slots=True, order=True)
class TournamentMaxRoundsOverride(_Exception4):
    """``max_rounds`` for one tournament game ``(root_seed, k, shuffle_index, game_index)``."""

    root_seed: int
    k: int
    shuffle_index: int
    game_index: int
    max_rounds: int
    _key: ClassVar = ("root_seed", "k", "shuffle_index", "game_index")

    def __post_init__(self) -> None:
        self._validate()
        if self.k < 2:
            raise ValueError("k must be at least 2")


@dataclass(frozen=True, slots=True, order=True)
class H2HMaxRoundsOverride(_Exception4):
    """``max_rounds`` for one H2H attempt ``(root_seed, pair_id, order, attempt_index)``."""

    root_seed: int
    pair_id: int
    order: int
    attempt_index: int
    max_rounds: int
    _key: ClassVar = ("root_seed", "pair_id", "order", "attempt_index")

    def __post_init__(self) -> None:
        self._validate()
        if self.order not in (0, 1):
            raise ValueError("order must be 0 or 1")


@dataclass(frozen=True, slots=True)
class GameLimits:
    target_score: int
    max_rounds: int


def _index(records: object, what: str, label: str) -> dict[Coordinate, int]:
    if not isinstance(records, tuple):
        raise TypeError(f"{what} must be a tuple")
    table: dict[Coordinate, int] = {}
    for rec in records:
        if rec.coordinate in table:
            raise ValueError(f"{label} max-round overrides contain duplicate coordinates")
        table[rec.coordinate] = rec.max_rounds
    return table


@dataclass(frozen=True, slots=True)
class GameProfile:
    default_target_score: int = 10_000
    default_max_rounds: int = 200
    tournament_max_rounds_overrides: tuple[TournamentMaxRoundsOverride, ...] = ()
    h2h_max_rounds_overrides: tuple[H2HMaxRoundsOverride, ...] = ()
    _t_index: dict = field(default_factory=dict, init=False, repr=False, compare=False)
    _h_index: dict = field(default_factory=dict, init=False, repr=False, compare=False)

    def __post_init__(self) -> None:
        if not _is_count(self.default_target_score) or self.default_target_score == 0:
            raise ValueError("default_target_score must be a positive integer")
        _check_count(self.default_max_rounds, "max_rounds")
        object.__setattr__(self, "_t_index", _index(
            self.tournament_max_rounds_overrides, "tournament_max_rounds_overrides", "tournament"))
        object.__setattr__(self, "_h_index", _index(
            self.h2h_max_rounds_overrides, "h2h_max_rounds_overrides", "H2H"))

    # ---- identity -----------------------------------------------------------------------
    def canonical_payload(self) -> dict[str, object]:
        """Key-for-key the reference's payload (game_profile.py:107-128), so hashes agree."""
        def listed(records: tuple) -> list[dict[str, int]]:
            return [asdict(r) for r in sorted(records, key=lambda r: r.coordinate)]

        payload: dict[str, object] = {"game_profile_contract_version": GAME_PROFILE_CONTRACT_VERSION}
        for f in fields(self):
            if f.name.startswith("_"):
                continue
            value = getattr(self, f.name)
            payload[f.name] = listed(value) if isinstance(value, tuple) else value
        return payload

    @property
    def sha256(self) -> str:
        """SHA-256 of the canonical JSON text (utils/authenticated_contract.py:99-114)."""
        text = json.dumps(self.canonical_payload(), sort_keys=True, separators=(",", ":"),
                          ensure_ascii=False, allow_nan=False)
        return hashlib.sha256(text.encode("utf-8")).hexdigest()

    # ---- single-game views (what the reference's callers ask) ------------------------------
    def _limits(self, index: dict[Coordinate, int], coordinate: Coordinate) -> GameLimits:
        return GameLimits(self.default_target_score,
                          index.get(coordinate, self.default_max_rounds))

    def tournament_limits(self, *, root_seed: int, k: int, shuffle_index: int,
                          game_index: int) -> GameLimits:
        return self._limits(self._t_index, (root_seed, k, shuffle_index, game_index))

    def h2h_limits(self, *, root_seed: int, pair_id: int, order: int,
                   attempt_index: int) -> GameLimits:
        return self._limits(self._h_index, (root_seed, pair_id, order, attempt_index))

    # ---- launch-side view ---------------------------------------------------------------
    def tournament_overrides_for(self, root_seed: int, k: int, shuffle0: int, n_shuffles: int
                                 ) -> list[tuple[int, int, int]]:
        """``(shuffle_index, game_index, max_rounds)`` rows inside one launch's shuffle range."""
        stop = shuffle0 + n_shuffles
        return [(sh, game, rounds) for (root, kk, sh, game), rounds in self._t_index.items()
                if root == root_seed and kk == k and shuffle0 <= sh < stop]


__all__ = ["GAME_PROFILE_CONTRACT_VERSION", "GameLimits", "GameProfile", "H2HMaxRoundsOverride",
           "TournamentMaxRoundsOverride"]
