"""Limit-only game settings (target score, safety-round limit, per-coordinate overrides).

Mirror of the reference's ``farkle.simulation.game_profile`` (src/farkle/simulation/
game_profile.py:24-200): same dataclasses, same validation, same canonical identity hash.
The kernels take ``default_target_score`` / ``default_max_rounds`` as launch scalars and the
``max_rounds`` overrides as a small keyed table (``fb_play_tournament`` override_* arguments).
"""

from __future__ import annotations

import hashlib
import json
from dataclasses import asdict, dataclass

GAME_PROFILE_CONTRACT_VERSION = 1


def _require_coordinate(value: int, *, name: str) -> None:
    if isinstance(value, bool) or not isinstance(value, int) or value < 0:
        raise ValueError(f"{name} must be a non-negative integer")


def _require_max_rounds(value: int) -> None:
    if isinstance(value, bool) or not isinstance(value, int) or value < 0:
        raise ValueError("max_rounds must be a non-negative integer")


@dataclass(frozen=True, slots=True, order=True)
class TournamentMaxRoundsOverride:
    root_seed: int
    k: int
    shuffle_index: int
    game_index: int
    max_rounds: int

    def __post_init__(self) -> None:
        for name in ("root_seed", "k", "shuffle_index", "game_index"):
            _require_coordinate(getattr(self, name), name=name)
        if self.k < 2:
            raise ValueError("k must be at least 2")
        _require_max_rounds(self.max_rounds)

    @property
    def coordinate(self) -> tuple[int, int, int, int]:
        return (self.root_seed, self.k, self.shuffle_index, self.game_index)


@dataclass(frozen=True, slots=True, order=True)
class H2HMaxRoundsOverride:
    root_seed: int
    pair_id: int
    order: int
    attempt_index: int
    max_rounds: int

    def __post_init__(self) -> None:
        for name in ("root_seed", "pair_id", "order", "attempt_index"):
            _require_coordinate(getattr(self, name), name=name)
        if self.order not in (0, 1):
            raise ValueError("order must be 0 or 1")
        _require_max_rounds(self.max_rounds)

    @property
    def coordinate(self) -> tuple[int, int, int, int]:
        return (self.root_seed, self.pair_id, self.order, self.attempt_index)


@dataclass(frozen=True, slots=True)
class GameLimits:
    target_score: int
    max_rounds: int


@dataclass(frozen=True, slots=True)
class GameProfile:
    default_target_score: int = 10_000
    default_max_rounds: int = 200
    tournament_max_rounds_overrides: tuple[TournamentMaxRoundsOverride, ...] = ()
    h2h_max_rounds_overrides: tuple[H2HMaxRoundsOverride, ...] = ()

    def __post_init__(self) -> None:
        if (isinstance(self.default_target_score, bool)
                or not isinstance(self.default_target_score, int)
                or self.default_target_score <= 0):
            raise ValueError("default_target_score must be a positive integer")
        _require_max_rounds(self.default_max_rounds)
        if not isinstance(self.tournament_max_rounds_overrides, tuple):
            raise TypeError("tournament_max_rounds_overrides must be a tuple")
        if not isinstance(self.h2h_max_rounds_overrides, tuple):
            raise TypeError("h2h_max_rounds_overrides must be a tuple")
        t = [o.coordinate for o in self.tournament_max_rounds_overrides]
        if len(set(t)) != len(t):
            raise ValueError("tournament max-round overrides contain duplicate coordinates")
        h = [o.coordinate for o in self.h2h_max_rounds_overrides]
        if len(set(h)) != len(h):
            raise ValueError("H2H max-round overrides contain duplicate coordinates")

    def canonical_payload(self) -> dict[str, object]:
        return {
            "game_profile_contract_version": GAME_PROFILE_CONTRACT_VERSION,
            "default_target_score": self.default_target_score,
            "default_max_rounds": self.default_max_rounds,
            "tournament_max_rounds_overrides": [
                asdict(o) for o in sorted(self.tournament_max_rounds_overrides,
                                          key=lambda item: item.coordinate)],
            "h2h_max_rounds_overrides": [
                asdict(o) for o in sorted(self.h2h_max_rounds_overrides,
                                          key=lambda item: item.coordinate)],
        }

    @property
    def sha256(self) -> str:
        """Canonical-JSON SHA-256 (utils/authenticated_contract.py:99-114)."""
        data = json.dumps(self.canonical_payload(), sort_keys=True, separators=(",", ":"),
                          ensure_ascii=False, allow_nan=False).encode("utf-8")
        return hashlib.sha256(data).hexdigest()

    def tournament_limits(self, *, root_seed: int, k: int, shuffle_index: int,
                          game_index: int) -> GameLimits:
        coordinate = (root_seed, k, shuffle_index, game_index)
        max_rounds = self.default_max_rounds
        for o in self.tournament_max_rounds_overrides:
            if o.coordinate == coordinate:
                max_rounds = o.max_rounds
                break
        return GameLimits(target_score=self.default_target_score, max_rounds=max_rounds)

    def h2h_limits(self, *, root_seed: int, pair_id: int, order: int,
                   attempt_index: int) -> GameLimits:
        coordinate = (root_seed, pair_id, order, attempt_index)
        max_rounds = self.default_max_rounds
        for o in self.h2h_max_rounds_overrides:
            if o.coordinate == coordinate:
                max_rounds = o.max_rounds
                break
        return GameLimits(target_score=self.default_target_score, max_rounds=max_rounds)

    # ---- launch-side views -------------------------------------------------------------
    def tournament_overrides_for(self, root_seed: int, k: int, shuffle0: int, n_shuffles: int
                                 ) -> list[tuple[int, int, int]]:
        """``(shuffle_index, game_index, max_rounds)`` rows inside one launch's shuffle range."""
        return [(o.shuffle_index, o.game_index, o.max_rounds)
                for o in self.tournament_max_rounds_overrides
                if o.root_seed == root_seed and o.k == k
                and shuffle0 <= o.shuffle_index < shuffle0 + n_shuffles]


__all__ = ["GAME_PROFILE_CONTRACT_VERSION", "GameLimits", "GameProfile", "H2HMaxRoundsOverride",
           "TournamentMaxRoundsOverride"]
