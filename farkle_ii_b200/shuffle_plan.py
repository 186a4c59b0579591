"""How many shuffles a (root, k) cell needs -- i.e. how large the launches of this engine are.

Stands in for the reference's ``farkle.simulation.workload_planner`` (src/farkle/simulation/
workload_planner.py:18-193) with the same public names and numbers: the widest Wilson interval a
binomial sample of n shuffles can have, the smallest n that brings it under a resolution, and the
plan record (``batch_count`` equal contiguous batches, never fewer than ``min_shuffles_per_batch``
each).  Pure host float arithmetic; the normal quantile is ``scipy.special.ndtri`` (the function
behind the reference's ``norm.ppf``, so every float in the plan has the same bits), and the plan
also reports the launch geometry it implies on a B200 (games per launch and seat-record bytes)
through ``launch_footprint``.
"""

from __future__ import annotations

import dataclasses
import math

try:                                            # scipy's ndtri is what norm.ppf evaluates: same bits
    from scipy.special import ndtri as _ndtri
except ImportError:                             # pragma: no cover - scipy is in the image
    from statistics import NormalDist

    def _ndtri(p: float) -> float:              # agrees to ~2 ulp; widths may differ in the last bit
        return NormalDist().inv_cdf(p)

WORKLOAD_PLAN_VERSION = 1
CAP_CONFIG_KEY = "screening.max_shuffles_per_root_k"
_SEAT_RECORD_BYTES = 80          # csrc/play.cuh: struct Seat


def _quantile(confidence: float) -> float:
    if not 0.0 < confidence < 1.0:
        raise ValueError("confidence must be between 0 and 1")
    return float(_ndtri(0.5 + confidence / 2.0))


def _whole(value: object, least: int) -> bool:
    return type(value) is not bool and isinstance(value, int) and value >= least


def worst_case_wilson_width(n: int, *, confidence: float = 0.95) -> float:
    """Widest full Wilson interval over all outcomes of n trials (workload_planner.py:72-90).

    The width peaks at the outcome(s) nearest p = 1/2, so only floor(n/2) and ceil(n/2) are tried.
    """
    if not _whole(n, 1):
        raise ValueError("n must be a positive integer")
    z = _quantile(confidence)
    shrink = 1.0 + z * z / n
    widest = 0.0
    for hits in {n // 2, n - n // 2}:
        p = hits / n
        half = z * math.sqrt(p * (1.0 - p) / n + z * z / (4.0 * n * n))
        widest = max(widest, 2.0 * half / shrink)
    return widest


def minimum_shuffles_for_resolution(resolution_delta: float, *, confidence: float = 0.95) -> int:
    """Smallest n with ``worst_case_wilson_width(n) <= resolution_delta`` (workload_planner.py:93-119)."""
    if not 0.0 < resolution_delta < 1.0:
        raise ValueError("resolution_delta must be between 0 and 1")
    _quantile(confidence)

    def meets(n: int) -> bool:
        return worst_case_wilson_width(n, confidence=confidence) <= resolution_delta

    hi = 1
    while not meets(hi):                    # gallop to an upper bracket
        hi *= 2
    lo = hi // 2                            # lo fails (or is 0), hi meets
    while hi - lo > 1:
        mid = (lo + hi) // 2
        lo, hi = (lo, mid) if meets(mid) else (mid, hi)
    return hi


@dataclasses.dataclass(frozen=True, slots=True)
class TournamentWorkloadPlan:
    root_seed: int
    k: int
    strategy_count: int
    confidence: float
    resolution_delta: float
    required_shuffles_unrounded: int
    required_shuffles: int
    batch_count: int
    shuffles_per_batch: int
    batch_construction: str
    games_per_shuffle: int
    required_games: int
    achieved_resolution: float
    shuffle_cap: int | None
    cap_exceeded: bool
    achieved_resolution_at_cap: float | None
    projected_games_per_second: float | None = None
    projected_runtime_seconds: float | None = None
    plan_version: int = WORKLOAD_PLAN_VERSION

    @property
    def status(self) -> str:
        return ("not_started", "blocked_by_cap")[bool(self.cap_exceeded)]

    def with_games_per_second(self, games_per_second: float) -> "TournamentWorkloadPlan":
        rate = float(games_per_second)
        if not (math.isfinite(rate) and rate > 0.0):
            raise ValueError("games_per_second must be finite and positive")
        return dataclasses.replace(self, projected_games_per_second=rate,
                                   projected_runtime_seconds=self.required_games / rate)

    def to_dict(self) -> dict[str, object]:
        out = dataclasses.asdict(self)
        out.update(status=self.status, cap_config_key=CAP_CONFIG_KEY)
        return out

    def launch_footprint(self) -> dict[str, int]:
        """Games and seat-record bytes of one batch launch and of the whole cell on the device."""
        per_batch = self.shuffles_per_batch * self.games_per_shuffle
        return {"games_per_batch_launch": per_batch,
                "seat_bytes_per_batch_launch": per_batch * self.k * _SEAT_RECORD_BYTES,
                "games_per_cell_launch": self.required_games,
                "seat_bytes_per_cell_launch": self.required_games * self.k * _SEAT_RECORD_BYTES}


class WorkloadCapExceeded(RuntimeError):
    """The plan needs more shuffles than the configured cap allows."""

    def __init__(self, plan: TournamentWorkloadPlan) -> None:
        self.plan = plan
        need = plan.required_shuffles
        super().__init__(f"Required {need} shuffles for root={plan.root_seed}, k={plan.k}, but "
                         f"{CAP_CONFIG_KEY}={plan.shuffle_cap}. Raise {CAP_CONFIG_KEY} to at "
                         f"least {need} and resume.")


def plan_tournament_workload(*, root_seed: int, k: int, strategy_count: int, resolution_delta: float,
                             confidence: float = 0.95, batch_count: int = 100,
                             min_shuffles_per_batch: int = 30, shuffle_cap: int | None = None,
                             projected_games_per_second: float | None = None) -> TournamentWorkloadPlan:
    """Precision, batches, game count and cap state of one cell (workload_planner.py:122-193)."""
    if not _whole(k, 2):
        raise ValueError("k must be an integer of at least 2")
    if not _whole(strategy_count, k) or strategy_count % k:
        raise ValueError("strategy_count must be a positive multiple of k")
    if not _whole(batch_count, 2):
        raise ValueError("batch_count must be an integer of at least 2")
    if not _whole(min_shuffles_per_batch, 1):
        raise ValueError("min_shuffles_per_batch must be a positive integer")
    if shuffle_cap is not None and not _whole(shuffle_cap, 1):
        raise ValueError("shuffle_cap must be positive when configured")

    bare = minimum_shuffles_for_resolution(resolution_delta, confidence=confidence)
    per_batch = max(min_shuffles_per_batch, -(-bare // batch_count))
    shuffles = per_batch * batch_count
    games_per_shuffle = strategy_count // k
    over = shuffle_cap is not None and shuffles > shuffle_cap
    plan = TournamentWorkloadPlan(
        root_seed=int(root_seed), k=k, strategy_count=strategy_count,
        confidence=float(confidence), resolution_delta=float(resolution_delta),
        required_shuffles_unrounded=bare, required_shuffles=shuffles,
        batch_count=batch_count, shuffles_per_batch=per_batch,
        batch_construction="equal_contiguous", games_per_shuffle=games_per_shuffle,
        required_games=shuffles * games_per_shuffle,
        achieved_resolution=worst_case_wilson_width(shuffles, confidence=confidence),
        shuffle_cap=shuffle_cap, cap_exceeded=over,
        achieved_resolution_at_cap=(worst_case_wilson_width(shuffle_cap, confidence=confidence)
                                    if over else None))
    if projected_games_per_second is None:
        return plan
    return plan.with_games_per_second(projected_games_per_second)


__all__ = ["CAP_CONFIG_KEY", "WORKLOAD_PLAN_VERSION", "TournamentWorkloadPlan", "WorkloadCapExceeded",
           "minimum_shuffles_for_resolution", "plan_tournament_workload", "worst_case_wilson_width"]
