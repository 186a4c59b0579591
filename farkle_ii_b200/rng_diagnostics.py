"""RNG lag diagnostics (strategy and matchup groups) from the tournament reduction itself.

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

"Matchup" groups (the games that seat the same multiset of strategies, one n_rounds observation per
game) come from the same launch: the device sorts the games of the cell by matchup key, keeps the
groups that have enough games and returns their lag sums (csrc/matchup.cuh).  They need the whole
(root, k) cell in ONE launch -- a cell of the mega config is 4,300 shuffles, far below the 2^31
games a launch takes.  The reference's group identity (BLAKE2b digest), its deterministic priority
cap and the report rows are host arithmetic on those few groups: ``MatchupLagGroups``.
"""

from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np

from .layout import LAG_WIDTH, MATCHUP_LAG_WIDTH

EXPECTED_NOTE = (
    "Zero-centered approximate descriptive reference band only; values inside or "
    "outside the band do not establish or refute independence"
)
STRATEGY_SEQUENCE_ORDER = "root_seed,k,shuffle_index,game_index,seat_index"
MATCHUP_SEQUENCE_ORDER = "root_seed,k,shuffle_index,game_index"
DEFAULT_MAX_MATCHUP_GROUPS = 100_000       # analysis.rng_max_matchup_groups (config.py:333)
GROUP_STRATEGY, GROUP_MATCHUP = 0, 1
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
        return _pearson(int(row[0]), row[1 + 5 * metric:6 + 5 * metric])

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


# ---- matchup groups ------------------------------------------------------------------------------
def _pearson(pairs: int, sums: Sequence[Any]) -> tuple[float | None, str]:
    """``_OnlineMetric.result`` (:2066-2076) on integer sums, with the reference's float expressions."""
    if pairs < 2:
        return None, "insufficient_pairs"
    sx, sy, sx2, sy2, sxy = (np.float64(v) for v in sums)
    count = float(pairs)
    numerator = count * sxy - sx * sy
    den_x = count * sx2 - sx ** 2
    den_y = count * sy2 - sy ** 2
    if den_x <= 0.0 or den_y <= 0.0:
        return None, "zero_variance"
    return float(numerator / (den_x * den_y) ** 0.5), "estimated"


def _splitmix64(values: np.ndarray) -> np.ndarray:
    """SplitMix64 finaliser over a uint64 array (:1267-1272); wraps modulo 2^64."""
    v = values.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        v += np.uint64(0x9E3779B97F4A7C15)
        v = (v ^ (v >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        v = (v ^ (v >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return v ^ (v >> np.uint64(31))


def matchup_group_ids(k: int, participants: np.ndarray, max_players: int) -> np.ndarray:
    """The reference's 64-bit matchup ids (``_matchup_ids``, :1173-1184): BLAKE2b-8, personalised
    ``farkle-m``, over the little-endian int32 row ``[k, sorted ids..., -1 padding]`` whose width
    is one plus the number of seat-strategy columns of the combined table (``max_players``)."""
    out = np.empty(len(participants), dtype=np.uint64)
    row = np.full(1 + max_players, -1, dtype="<i4")
    row[0] = k
    for g, ids in enumerate(np.asarray(participants, dtype=np.int32)):
        row[1:1 + k] = ids
        digest = hashlib.blake2b(row.tobytes(), digest_size=8, person=b"farkle-m").digest()
        out[g] = int.from_bytes(digest, "little", signed=False)
    return out


@dataclass
class MatchupLagGroups:
    """Lag sums of the matchup groups of one (root, k) cell that have enough observations."""

    lags: tuple[int, ...]
    k: int
    participants: np.ndarray   # int32 [groups, k] sorted strategy ids
    counts: np.ndarray         # int64 [groups] observations (games)
    stats: np.ndarray          # int64 [groups, n_lags, MATCHUP_LAG_WIDTH]

    @classmethod
    def from_launch(cls, lags: Sequence[int], result: Any) -> "MatchupLagGroups":
        """From the ``matchup_*`` outputs of a ``play_tournament(..., matchup_min_observations=m)``
        launch that covered the whole cell; groups are put in participant order."""
        part = result.matchup_participants.cpu().numpy().astype(np.int32)
        counts = result.matchup_count.cpu().numpy().view(np.uint32).astype(np.int64)
        stats = result.matchup_stats.cpu().numpy().astype(np.int64)
        return cls(tuple(int(v) for v in lags), int(result.k), part, counts, stats)._canonical()

    @classmethod
    def from_rows(cls, lags: Sequence[int], rows: np.ndarray, min_observations: int) -> "MatchupLagGroups":
        """Brute force over compact rows in game order (host reference for the device path)."""
        lags = tuple(int(v) for v in lags)
        k = rows["seats"].shape[1]
        series: dict[tuple[int, ...], list[int]] = {}
        for row in rows:
            series.setdefault(tuple(sorted(int(v) for v in row["seats"]["strategy"])), []).append(
                int(row["n_rounds"]))
        keep = sorted(key for key, seq in series.items() if len(seq) >= min_observations)
        stats = np.zeros((len(keep), len(lags), MATCHUP_LAG_WIDTH), dtype=np.int64)
        for g, key in enumerate(keep):
            seq = series[key]
            for z, lag in enumerate(lags):
                for j in range(lag, len(seq)):
                    x, y = seq[j - lag], seq[j]
                    stats[g, z] += [1, x, y, x * x, y * y, x * y]
        return cls(lags, k, np.array(keep, dtype=np.int32).reshape(len(keep), k),
                   np.array([len(series[key]) for key in keep], dtype=np.int64), stats)

    def _canonical(self) -> "MatchupLagGroups":
        order = np.lexsort(tuple(self.participants[:, c] for c in reversed(range(self.k))))
        return MatchupLagGroups(self.lags, self.k, self.participants[order], self.counts[order],
                                self.stats[order])

    def __len__(self) -> int:
        return len(self.counts)

    def group_ids(self, max_players: int) -> np.ndarray:
        return matchup_group_ids(self.k, self.participants, max_players)

    def priorities(self, max_players: int) -> np.ndarray:
        """Stable selection priority of every group (``_priority``, :1281-1284)."""
        values = self.group_ids(max_players) ^ (np.uint64(self.k) << np.uint64(48))
        values ^= np.uint64(GROUP_MATCHUP) << np.uint64(63)
        return _splitmix64(values ^ np.uint64(0xD1B54A32D192ED03))

    def rows(self, max_players: int, keep: np.ndarray | None = None) -> list[dict[str, Any]]:
        """Report rows of the (selected) groups as ``_rows_for_online_group`` builds them for a
        matchup (:2110-2159): n_rounds only, one row per lag; groups in (group id, ids) order."""
        ids = self.group_ids(max_players)
        chosen = np.flatnonzero(np.ones(len(self), dtype=bool) if keep is None else keep)
        order = chosen[np.lexsort((*(self.participants[chosen, c] for c in reversed(range(self.k))),
                                   ids[chosen]))]
        out: list[dict[str, Any]] = []
        for g in order:
            people = [int(v) for v in self.participants[g]]
            for z, lag in enumerate(self.lags):
                pairs = int(self.stats[g, z, 0])
                value, status = _pearson(pairs, self.stats[g, z, 1:6])
                half_width = 1.96 / pairs**0.5 if pairs > 0 else None
                out.append({
                    "summary_level": "matchup",
                    "strategy": None,
                    "matchup_id": int(ids[g]),
                    "matchup": " | ".join(str(v) for v in people),
                    "participant_strategy_ids": people,
                    "n_players": self.k,
                    "observations": int(self.counts[g]),
                    "lagged_pairs": pairs,
                    "lag": lag,
                    "metric": "n_rounds",
                    "autocorr": value,
                    "estimability_status": status,
                    "zero_centered_descriptive_reference_band_lower": (
                        -half_width if half_width is not None else None),
                    "zero_centered_descriptive_reference_band_upper": half_width,
                    "sequence_order": MATCHUP_SEQUENCE_ORDER,
                    "note": EXPECTED_NOTE,
                })
        return out


def select_matchup_groups(cells: Sequence[MatchupLagGroups], max_players: int,
                          cap: int | None = DEFAULT_MAX_MATCHUP_GROUPS) -> list[np.ndarray]:
    """The reference's deterministic cap over the eligible matchup groups of ALL player counts
    (``_write_or_reuse_selection``, :1611-1651): keep the ``cap`` groups that come first by
    ``(priority, group_type, k, group_id, p0, p1, ...)``.  Returns one keep-mask per cell."""
    total = sum(len(c) for c in cells)
    if cap is None or total <= cap:
        return [np.ones(len(c), dtype=bool) for c in cells]
    keys = np.zeros((total, 4 + max_players), dtype=np.uint64)     # priority, k, group id, p..., owner
    owner = np.empty(total, dtype=np.int64)
    at = 0
    for index, cell in enumerate(cells):
        n = len(cell)
        keys[at:at + n, 0] = cell.priorities(max_players)
        keys[at:at + n, 1] = cell.k
        keys[at:at + n, 2] = cell.group_ids(max_players)
        padded = np.full((n, max_players), -1, dtype=np.int64)
        padded[:, :cell.k] = cell.participants
        keys[at:at + n, 3:3 + max_players] = (padded + 2**31).astype(np.uint64)   # order-preserving
        owner[at:at + n] = index
        at += n
    order = np.lexsort(tuple(keys[:, c] for c in reversed(range(3 + max_players))))[:cap]
    masks = [np.zeros(len(c), dtype=bool) for c in cells]
    starts = np.cumsum([0, *(len(c) for c in cells)])
    for pos in order:
        masks[owner[pos]][pos - starts[owner[pos]]] = True
    return masks


def diagnose_root(root_seed: int, ks: Sequence[int], n_shuffles: int, strategies: Any,
                  strategy_ids: Sequence[int], lags: Sequence[int] | None = None, *, max_players: int,
                  cap: int | None = DEFAULT_MAX_MATCHUP_GROUPS, engine: Any = None,
                  device: int | None = None, **limits: int) -> list[dict[str, Any]]:
    """Every row of the reference's ``rng_diagnostics.parquet`` for one root: each (root, k) cell is
    played once with the lag outputs switched on; strategy groups of every k, then the matchup
    groups that survive eligibility (``min(lags) + 2`` games) and the priority cap across all k.
    ``max_players`` is the number of seat-strategy columns of the combined table (the largest k
    of the run); row order is (level, k, group) -- the reference's is by hash partition."""
    from . import device as fdev

    lags = normalize_lags(lags)
    eng = engine if engine is not None else fdev.get_engine(device)
    ids = np.asarray(strategy_ids, dtype=np.int32)
    rows: list[dict[str, Any]] = []
    cells: list[MatchupLagGroups] = []
    for k in ks:
        res = eng.play_tournament(root_seed, k, 0, n_shuffles, strategies, strategy_ids=ids, lags=lags,
                                  matchup_min_observations=minimum_observations(lags), **limits)
        rows += StrategyLagState.from_launch(lags, n_shuffles, res.lag_stats, res.lag_edges).rows(ids.tolist(), k)
        cells.append(MatchupLagGroups.from_launch(lags, res))
    for cell, keep in zip(cells, select_matchup_groups(cells, max_players, cap)):
        rows += cell.rows(max_players, keep)
    return rows


def diagnostics_schema():
    """Arrow schema of ``rng_diagnostics.parquet`` (``_stats_schema``, :2079-2098)."""
    import pyarrow as pa

    text, f64 = pa.string(), pa.float64()
    return pa.schema([
        pa.field("summary_level", text, nullable=False), pa.field("strategy", pa.int32()),
        pa.field("matchup_id", pa.uint64()), pa.field("matchup", text),
        pa.field("participant_strategy_ids", pa.list_(pa.int32())),
        pa.field("n_players", pa.int16(), nullable=False),
        pa.field("observations", pa.int64(), nullable=False),
        pa.field("lagged_pairs", pa.int64(), nullable=False), pa.field("lag", pa.int32(), nullable=False),
        pa.field("metric", text, nullable=False), pa.field("autocorr", f64),
        pa.field("estimability_status", text, nullable=False),
        pa.field("zero_centered_descriptive_reference_band_lower", f64),
        pa.field("zero_centered_descriptive_reference_band_upper", f64),
        pa.field("sequence_order", text, nullable=False), pa.field("note", text, nullable=False),
    ])


def diagnostics_table(rows: Sequence[dict[str, Any]]):
    """Report rows (``StrategyLagState.rows`` / ``MatchupLagGroups.rows``) as the Arrow table the
    reference writes (``pa.Table.from_pylist(rows, schema=_stats_schema())``, :2162-2207)."""
    import pyarrow as pa

    return pa.Table.from_pylist(list(rows), schema=diagnostics_schema())


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


def join_lag_segments(segments: Sequence[tuple[int, StrategyLagState]]) -> StrategyLagState:
    """Join ``(first shuffle, state)`` segments that together tile a run of shuffles."""
    ordered = sorted(segments, key=lambda seg: seg[0])
    if not ordered:
        raise ValueError("no lag segments to join")
    at, joined = ordered[0][0] + ordered[0][1].n_obs, ordered[0][1]
    for s0, part in ordered[1:]:
        if s0 != at:
            raise ValueError(f"lag segments do not tile the shuffle range: expected {at}, got {s0}")
        joined = joined.extend(part)
        at += part.n_obs
    return joined


def gather_lag_segments(segments: Sequence[tuple[int, StrategyLagState]]) -> StrategyLagState:
    """All ranks' segments joined in shuffle order; every rank gets the result.  Payload per
    segment: the sums and 2 x max_lag observations per strategy (the path's tallies travel by
    all-reduce; these small states by one all-gather)."""
    import torch.distributed as dist

    everything = list(segments)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts: list[Any] = [None] * dist.get_world_size()
        dist.all_gather_object(parts, list(segments))
        everything = [seg for part in parts for seg in part]
    return join_lag_segments(everything)


def cell_lag_state(root_seed: int, k: int, num_shuffles: int, strategies: Any,
                   lags: Sequence[int] | None = None, *, batch_size: int, rank: int = 0, world: int = 1,
                   engine: Any = None, device: int | None = None, **limits: int) -> StrategyLagState:
    """Strategy-group lag state of a whole (root, k) cell played the way ``run_cell`` shards it:
    deterministic batches dealt in contiguous blocks over the ranks, one launch per contiguous run, the
    per-launch states gathered and joined in shuffle order."""
    from . import device as fdev
    from .run_tournament import merge_ranges, shard_batches

    lags = normalize_lags(lags)
    eng = engine if engine is not None else fdev.get_engine(device)
    segments = []
    for s0, cnt in merge_ranges(shard_batches(num_shuffles, batch_size, rank, world)):
        res = eng.play_tournament(root_seed, k, s0, cnt, strategies, lags=lags, **limits)
        segments.append((s0, StrategyLagState.from_launch(lags, cnt, res.lag_stats, res.lag_edges)))
    return gather_lag_segments(segments)


__all__ = ["DEFAULT_MAX_MATCHUP_GROUPS", "EXPECTED_NOTE", "MATCHUP_SEQUENCE_ORDER", "MatchupLagGroups",
           "STRATEGY_SEQUENCE_ORDER", "StrategyLagState", "cell_lag_state", "diagnostics_schema",
           "diagnostics_table", "diagnose_root", "gather_lag_segments",
           "join_lag_segments", "matchup_group_ids",
           "minimum_observations", "normalize_lags", "observations_from_rows", "select_matchup_groups",
           "strategy_lag_state"]
