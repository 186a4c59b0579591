"""RNG lag diagnostics of the strategy groups, from the tournament reduction itself.

The reference stage (src/farkle/analysis/rng_diagnostics.py) re-reads every curated row, routes and
externally sorts one observation per seat exposure by ``(group, root_seed, shuffle_index,
game_index, seat_index)`` and pushes ``win_indicator`` / ``n_rounds`` through an online lagged-pair
accumulator (``_OnlineMetric``, :2032-2077) to report a Pearson autocorrelation per lag
(``_rows_for_online_group``, :2110-2159).  For "strategy" groups none of that data movement is
needed here: a strategy is seated exactly once per shuffle, so its sequence inside a (root, k) cell
is "its game in shuffle 0, 1, 2, ...", which ``lag_gather_kernel`` (csrc/play.cuh) walks through
the inverse permutations right after the games were played.  The six sums per (strategy, lag,
metric) are integers, so the device's int64 totals equal the reference's float64 accumulators
exactly and the final floats below are computed with the same expressions.

A cell may be played in several launches (deterministic batches, ranks): ``StrategyLagState`` is a
monoid -- ``a.extend(b)`` adds the pairs that straddle the boundary from the ``max(lags)``
observations each launch reports at its two ends.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np

from .layout import LAG_WIDTH

EXPECTED_NOTE = (
    "Zero-centered approximate descriptive reference band only; values inside or "
    "outside the band do not establish or refute independence"
)
STRATEGY_SEQUENCE_ORDER = "root_seed,k,shuffle_index,game_index,seat_index"
WIN_BIT = 1 << 16           # observation word: n_rounds | win << 16 (include/farkle_b200.h)


def normalize_lags(lags: Sequence[int] | None) -> tuple[int, ...]:
    """Distinct positive lags in ascending order; ``None`` means ``(1,)`` (:972-975)."""
    if lags is None:
        return (1,)
    return tuple(sorted({int(v) for v in lags if int(v) > 0}))


def minimum_observations(lags: Sequence[int]) -> int:
    """Groups with fewer observations are not reported (:592)."""
    return min(lags) + 2


def _pair_sums(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """[n, LAG_WIDTH] contribution of one (earlier, current) observation column pair."""
    xw, yw = (x >> 16).astype(np.int64), (y >> 16).astype(np.int64)
    xr, yr = (x & 0xFFFF).astype(np.int64), (y & 0xFFFF).astype(np.int64)
    one = np.ones_like(xw)
    return np.stack([one, xw, yw, xw, yw, xw * yw, xr, yr, xr * xr, yr * yr, xr * yr], axis=1)


@dataclass
class StrategyLagState:
    """Lag sums of every table entry over a contiguous run of shuffles of one (root, k) cell."""

    lags: tuple[int, ...]
    n_obs: int                 # shuffles covered (= observations per strategy)
    stats: np.ndarray          # int64 [n_strategies, n_lags, LAG_WIDTH]
    head: np.ndarray           # uint32 [n_strategies, min(max_lag, n_obs)] first observations
    tail: np.ndarray           # uint32 [n_strategies, min(max_lag, n_obs)] last observations

    @classmethod
    def empty(cls, n_strategies: int, lags: Sequence[int]) -> "StrategyLagState":
        lags = tuple(int(v) for v in lags)
        z = np.zeros((n_strategies, 0), dtype=np.uint32)
        return cls(lags, 0, np.zeros((n_strategies, len(lags), LAG_WIDTH), dtype=np.int64), z, z.copy())

    @classmethod
    def from_launch(cls, lags: Sequence[int], n_shuffles: int, lag_stats: Any, lag_edges: Any
                    ) -> "StrategyLagState":
        """From the ``lag_stats`` / ``lag_edges`` outputs of one ``play_tournament`` launch."""
        lags = tuple(int(v) for v in lags)
        stats = np.asarray(lag_stats.cpu() if hasattr(lag_stats, "cpu") else lag_stats, dtype=np.int64)
        edges = np.asarray(lag_edges.cpu() if hasattr(lag_edges, "cpu") else lag_edges).view(np.uint32)
        h = min(max(lags), n_shuffles)
        return cls(lags, int(n_shuffles), stats.copy(), edges[:, 0, :h].copy(), edges[:, 1, :h].copy())

    @classmethod
    def from_observations(cls, lags: Sequence[int], obs: np.ndarray) -> "StrategyLagState":
        """Brute force from a full ``[n_strategies, n_obs]`` observation matrix (host reference)."""
        lags = tuple(int(v) for v in lags)
        obs = np.ascontiguousarray(obs, dtype=np.uint32)
        n, m = obs.shape
        stats = np.zeros((n, len(lags), LAG_WIDTH), dtype=np.int64)
        for z, lag in enumerate(lags):
            for j in range(lag, m):
                stats[:, z] += _pair_sums(obs[:, j - lag], obs[:, j])
        h = min(max(lags), m)
        return cls(lags, m, stats, obs[:, :h].copy(), obs[:, m - h:].copy())

    @property
    def max_lag(self) -> int:
        return max(self.lags)

    def extend(self, other: "StrategyLagState") -> "StrategyLagState":
        """State of this run of shuffles followed immediately by ``other``'s."""
        if self.lags != other.lags or self.stats.shape != other.stats.shape:
            raise ValueError("lag states of different lags / strategy tables cannot be joined")
        stats = self.stats + other.stats
        ha = self.tail.shape[1]
        for z, lag in enumerate(self.lags):
            for t in range(min(lag, other.n_obs)):
                at = ha - lag + t          # position of the earlier observation inside self.tail
                if at >= 0:
                    stats[:, z] += _pair_sums(self.tail[:, at], other.head[:, t])
        total = self.n_obs + other.n_obs
        h = min(self.max_lag, total)
        head = np.concatenate([self.head, other.head], axis=1)[:, :h]
        tail = np.concatenate([self.tail, other.tail], axis=1)
        return StrategyLagState(self.lags, total, stats, head.copy(), tail[:, tail.shape[1] - h:].copy())

    # ---- reporting (the reference's row dicts) ---------------------------------------------
    def autocorr(self, index: int, z: int, metric: int) -> tuple[float | None, str]:
        """``_OnlineMetric.result`` (:2066-2076) for table entry ``index``, lag slot ``z``;
        metric 0 = win_indicator, 1 = n_rounds.  Same float expressions, so the same bits."""
        row = self.stats[index, z]
        pairs = int(row[0])
        if pairs < 2:
            return None, "insufficient_pairs"
        sx, sy, sx2, sy2, sxy = (np.float64(v) for v in row[1 + 5 * metric:6 + 5 * metric])
        count = float(pairs)
        numerator = count * sxy - sx * sy
        den_x = count * sx2 - sx ** 2
        den_y = count * sy2 - sy ** 2
        if den_x <= 0.0 or den_y <= 0.0:
            return None, "zero_variance"
        return float(numerator / (den_x * den_y) ** 0.5), "estimated"

    def rows(self, strategy_ids: Sequence[int], k: int) -> list[dict[str, Any]]:
        """Rows of the diagnostics table for the strategy groups, as ``_rows_for_online_group``
        builds them (:2110-2159): per strategy win_indicator rows, then n_rounds rows, each over
        the lags.  Groups below ``minimum_observations`` are left out, like the reference's
        eligibility phase does."""
        if self.n_obs < minimum_observations(self.lags):
            return []
        out: list[dict[str, Any]] = []
        for index, sid in enumerate(strategy_ids):
            for metric, name in enumerate(("win_indicator", "n_rounds")):
                for z, lag in enumerate(self.lags):
                    value, status = self.autocorr(index, z, metric)
                    pairs = int(self.stats[index, z, 0])
                    half_width = 1.96 / pairs**0.5 if pairs > 0 else None
                    out.append({
                        "summary_level": "strategy",
                        "strategy": int(sid),
                        "matchup_id": None,
                        "matchup": None,
                        "participant_strategy_ids": None,
                        "n_players": k,
                        "observations": self.n_obs,
                        "lagged_pairs": pairs,
                        "lag": lag,
                        "metric": name,
                        "autocorr": value,
                        "estimability_status": status,
                        "zero_centered_descriptive_reference_band_lower": (
                            -half_width if half_width is not None else None),
                        "zero_centered_descriptive_reference_band_upper": half_width,
                        "sequence_order": STRATEGY_SEQUENCE_ORDER,
                        "note": EXPECTED_NOTE,
                    })
        return out


def observations_from_rows(rows: np.ndarray, n_strategies: int, n_shuffles: int) -> np.ndarray:
    """``[n_strategies, n_shuffles]`` observation words rebuilt from compact rows whose seat
    ``strategy`` field holds the TABLE position (no strategy_ids); used to check the kernel."""
    k = rows["seats"].shape[1]
    gps = n_strategies // k
    assert len(rows) == n_shuffles * gps
    shuffle = np.repeat(np.arange(n_shuffles), gps)
    obs = np.zeros((n_strategies, n_shuffles), dtype=np.uint32)
    safety = (rows["flags"] & 1) != 0
    for s in range(k):
        won = (~safety) & (rows["winner_seat"] == s)
        obs[rows["seats"]["strategy"][:, s], shuffle] = rows["n_rounds"].astype(np.uint32) | (
            won.astype(np.uint32) << 16)
    return obs


def strategy_lag_state(root_seed: int, k: int, shuffle0: int, n_shuffles: int, strategies: Any,
                       lags: Sequence[int] | None = None, *, batch_shuffles: int | None = None,
                       target_score: int = 10_000, max_rounds: int = 200, device: int | None = None,
                       engine: Any = None) -> StrategyLagState:
    """Play shuffles ``shuffle0 .. shuffle0 + n_shuffles - 1`` of cell (root_seed, k) -- in
    launches of ``batch_shuffles`` if given -- and return the joined lag state."""
    from . import device as fdev

    lags = normalize_lags(lags)
    eng = engine if engine is not None else fdev.get_engine(device)
    step = n_shuffles if not batch_shuffles else int(batch_shuffles)
    state: StrategyLagState | None = None
    for s0 in range(0, n_shuffles, max(step, 1)):
        cnt = min(step, n_shuffles - s0)
        res = eng.play_tournament(root_seed, k, shuffle0 + s0, cnt, strategies, target_score=target_score,
                                  max_rounds=max_rounds, lags=lags)
        part = StrategyLagState.from_launch(lags, cnt, res.lag_stats, res.lag_edges)
        state = part if state is None else state.extend(part)
    if state is None:
        n = len(strategies) if not hasattr(strategies, "numel") else strategies.numel() // 8
        state = StrategyLagState.empty(n, lags)
    return state


def gather_lag_states(state: StrategyLagState) -> StrategyLagState:
    """Join the states of all ranks in rank order (rank r holds the r-th contiguous shuffle
    range of the cell, as ``run_tournament.shard_batches`` deals them).  Every rank gets the
    result.  Payload per rank: the sums and 2 x max_lag observations per strategy."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return state
    parts: list[Any] = [None] * dist.get_world_size()
    dist.all_gather_object(parts, state)
    joined = parts[0]
    for part in parts[1:]:
        joined = joined.extend(part)
    return joined


__all__ = ["EXPECTED_NOTE", "STRATEGY_SEQUENCE_ORDER", "StrategyLagState", "gather_lag_states",
           "minimum_observations", "normalize_lags", "observations_from_rows", "strategy_lag_state"]
