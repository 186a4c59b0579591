"""Tournament host surface over the CUDA engine.

Mirror of the reference's ``farkle.simulation.run_tournament`` for the seams its tests and
callers already treat as replaceable:

* ``TournamentConfig`` / ``ShuffleTask``                run_tournament.py:78-104
* ``METRIC_LABELS``                                      :108-121
* ``OutcomeCounter``                                     :165-250
* ``_init_worker`` / ``WorkerState``                     :253-277
* ``_play_one_shuffle`` / ``_play_shuffle``              :301-400
* ``_run_chunk`` / ``_run_chunk_item`` / ``_run_chunk_metrics``   :403-585
* ``run_tournament``                                     :1050-1859 (reduce loop, checkpoint
  payload, ``{k}p_metrics.parquet``; the artifact-contract sidecars are out of scope)

Same names, same argument meaning, same return types.  The bodies differ: a chunk of
shuffles is ONE ``fb_play_tournament`` launch (contiguous shuffle runs), the per-strategy
tallies come back as an ``int64[ids][26]`` tensor and are unpacked into the reference's
``OutcomeCounter`` / ``dict[label][strategy] -> float`` shapes.  With ``torch.distributed``
initialised, ``run_tournament`` deals deterministic batches in contiguous blocks over ranks and
merges the tally tensors with one all-reduce per cell (NCCL on the GPUs).
"""

from __future__ import annotations

import json
import logging
import os
import pickle
import tempfile
from collections import Counter, defaultdict
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable, Dict, Iterator, List, Mapping, Sequence, Tuple

import numpy as np

from .limits import GameProfile
from .layout import N_METRICS, T_ATTEMPTED, T_COMPLETED, T_SAFETY, T_SQ_SUMS, T_SUMS, T_WINS, TALLY_WIDTH
from .random import RNG_SCHEME_VERSION, RandomPurpose, coordinate_seed
from .simulation import (
    OUTCOME_SCHEMA_VERSION,
    TOURNAMENT_METHOD_VERSION,
    _prepare_public_helper_strategies,
    check_row_flags,
    compact_rows_to_table,
    expand_rows,
)
from .strategies import ThresholdStrategy, generate_strategy_grid, pack_strategies

LOGGER = logging.getLogger(__name__)

NUM_SHUFFLES: int = 5_907
DESIRED_SEC_PER_CHUNK: int = 10
CKPT_EVERY_SEC: int = 30


@dataclass
class TournamentConfig:
    """Runtime configuration (run_tournament.py:78-94)."""

    n_players: int = 5
    num_shuffles: int = NUM_SHUFFLES
    desired_sec_per_chunk: int = DESIRED_SEC_PER_CHUNK
    ckpt_every_sec: int = CKPT_EVERY_SEC
    n_strategies: int = 7_140
    mp_start_method: str | None = None
    deterministic_batch_size: int = 30

    @property
    def games_per_shuffle(self) -> int:
        return self.n_strategies // self.n_players


@dataclass(frozen=True, slots=True)
class ShuffleTask:
    """Stable coordinate identity of one tournament shuffle (run_tournament.py:97-104)."""

    root_seed: int
    k: int
    shuffle_index: int
    shuffle_seed: int
    deterministic_batch_id: int


METRIC_LABELS: Tuple[str, ...] = (
    "winning_score", "n_rounds", "winner_farkles", "winner_rolls", "winner_highest_turn",
    "winner_smart_five_uses", "winner_n_smart_five_dice", "winner_smart_one_uses",
    "winner_n_smart_one_dice", "winner_hot_dice", "winner_hit_max_rounds",
)
assert len(METRIC_LABELS) == N_METRICS


class OutcomeCounter(Counter):
    """Win counter carrying attempted/completed exposure conservation (run_tournament.py:165-230)."""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        self.attempted_exposures: Counter = Counter()
        self.completed_exposures: Counter = Counter()
        self.safety_limit_exposures: Counter = Counter()
        self.games_attempted = 0
        self.games_completed = 0
        self.games_safety_limit = 0
        super().__init__(*args, **kwargs)

    def record_row(self, row: Mapping[str, Any], *, k: int, source: str) -> str:
        status = row.get("termination_status")
        if row.get("outcome_schema_version") != OUTCOME_SCHEMA_VERSION:
            raise RuntimeError(f"{source} is not outcome-schema-v{OUTCOME_SCHEMA_VERSION} compatible; "
                               "explicit-outcome tournament aggregation is disabled for this row")
        if status not in ("completed", "safety_limit"):
            raise RuntimeError(f"{source} has unsupported termination_status={status!r}")
        completed = status == "completed"
        has_winner = row.get("winner_seat") is not None or row.get("winner_strategy") is not None
        if not completed and has_winner:
            raise RuntimeError(f"{source} fabricates a winner for a safety-limit attempt")
        if completed and (row.get("winner_seat") is None or row.get("winner_strategy") is None):
            raise RuntimeError(f"{source} is completed but has no canonical winner")
        for seat in range(1, k + 1):
            strategy = row.get(f"P{seat}_strategy")
            if strategy is None:
                raise ValueError(f"{source} is missing strategy exposure for P{seat}")
            self.attempted_exposures[strategy] += 1
            (self.completed_exposures if completed else self.safety_limit_exposures)[strategy] += 1
        self.games_attempted += 1
        if completed:
            self.games_completed += 1
        else:
            self.games_safety_limit += 1
        return status

    def absorb(self, other: Counter) -> None:
        super().update(other)
        if isinstance(other, OutcomeCounter):
            self.attempted_exposures.update(other.attempted_exposures)
            self.completed_exposures.update(other.completed_exposures)
            self.safety_limit_exposures.update(other.safety_limit_exposures)
            self.games_attempted += other.games_attempted
            self.games_completed += other.games_completed
            self.games_safety_limit += other.games_safety_limit
            return
        completed = int(sum(other.values()))
        self.attempted_exposures.update(other)
        self.completed_exposures.update(other)
        self.games_attempted += completed
        self.games_completed += completed

    def outcome_payload(self) -> dict[str, Any]:
        return {
            "games_attempted": self.games_attempted,
            "games_completed": self.games_completed,
            "games_safety_limit": self.games_safety_limit,
            "attempted_exposures": dict(self.attempted_exposures),
            "completed_exposures": dict(self.completed_exposures),
            "safety_limit_exposures": dict(self.safety_limit_exposures),
        }

    def __reduce__(self):
        return (_restore_outcome_counter, (dict(self), self.outcome_payload()))


def _restore_outcome_counter(counts: dict, outcome_counts: dict[str, Any]) -> OutcomeCounter:
    restored = OutcomeCounter(counts)
    restored.attempted_exposures.update(outcome_counts.get("attempted_exposures", {}))
    restored.completed_exposures.update(outcome_counts.get("completed_exposures", {}))
    restored.safety_limit_exposures.update(outcome_counts.get("safety_limit_exposures", {}))
    restored.games_attempted = int(outcome_counts.get("games_attempted", 0))
    restored.games_completed = int(outcome_counts.get("games_completed", 0))
    restored.games_safety_limit = int(outcome_counts.get("games_safety_limit", 0))
    return restored


# --------------------------------------------------------------------------- tallies <-> counters
MetricSums = Dict[str, Dict[Any, float]]


def tallies_to_outcome(tallies: np.ndarray, ids: Sequence[int], games: Tuple[int, int, int],
                       first_seen: np.ndarray | None = None
                       ) -> Tuple[OutcomeCounter, MetricSums, MetricSums]:
    """``int64[n][26]`` -> ``(OutcomeCounter, sums, sq_sums)`` exactly as run_tournament.py:331-393
    would have built them: keys exist only where the reference would have touched them, and --
    given the launch's ``first_seen`` ordinals -- in the order the reference inserted them (first
    win / first exposure in game and seat order), so that pickles of the merged counters come
    out byte-identical.  Without ``first_seen`` keys are in table order."""
    t = np.asarray(tallies).reshape(-1, TALLY_WIDTH)
    seen = None
    if first_seen is not None:
        seen = np.asarray(first_seen).reshape(-1, 4)
        seen = seen.view(np.uint32) if seen.dtype == np.int32 else seen      # -1 = never sorts last

    def touched(col: int, seen_col: int) -> np.ndarray:
        rows = np.flatnonzero(t[:, col])
        return rows if seen is None else rows[np.argsort(seen[rows, seen_col], kind="stable")]

    wins = OutcomeCounter()
    sums: MetricSums = {m: defaultdict(float) for m in METRIC_LABELS}
    sq_sums: MetricSums = {m: defaultdict(float) for m in METRIC_LABELS}
    id_list = np.asarray(ids).tolist()           # python ints, as the reference's keys are

    def column(rows: np.ndarray, keys: list, col: int, as_float: bool = False) -> dict:
        vals = t[rows, col]
        return dict(zip(keys, (vals.astype(np.float64) if as_float else vals).tolist()))

    def keys_of(rows: np.ndarray) -> list:
        return [id_list[r] for r in rows.tolist()]

    for col, seen_col, target in ((T_ATTEMPTED, 1, wins.attempted_exposures),
                                  (T_COMPLETED, 2, wins.completed_exposures),
                                  (T_SAFETY, 3, wins.safety_limit_exposures)):
        rows = touched(col, seen_col)
        dict.update(target, column(rows, keys_of(rows), col))     # fresh Counter: plain insert, not add
    winners = touched(T_WINS, 0)
    win_keys = keys_of(winners)                  # one key list shared by the 23 winner dicts
    dict.update(wins, column(winners, win_keys, T_WINS))
    block = t[winners, T_SUMS:T_SUMS + 2 * len(METRIC_LABELS)].astype(np.float64).T.tolist()
    for m, label in enumerate(METRIC_LABELS):
        sums[label].update(zip(win_keys, block[m]))
        sq_sums[label].update(zip(win_keys, block[T_SQ_SUMS - T_SUMS + m]))
    wins.games_attempted, wins.games_completed, wins.games_safety_limit = (int(x) for x in games)
    return wins, sums, sq_sums


# --------------------------------------------------------------------------- worker state
@dataclass(slots=True)
class WorkerState:
    strats: list[ThresholdStrategy]
    cfg: TournamentConfig
    game_profile: GameProfile | None = None
    table: np.ndarray | None = None      # packed fb_strategy_t table
    ids: np.ndarray | None = None        # strategy id of table entry i
    device: int | None = None


_STATE: WorkerState | None = None


def _init_worker(strategies: Sequence[ThresholdStrategy], config: TournamentConfig,
                 game_profile: GameProfile | None = None, progress_endpoint: object | None = None,
                 *, device: int | None = None) -> None:
    """Initialise per-process state (run_tournament.py:265-277)."""
    global _STATE
    del progress_endpoint
    if len(strategies) % config.n_players != 0:
        raise ValueError(f"n_players must divide {len(strategies):,}")
    strats, table, ids = _prepared_table(strategies)
    config.n_strategies = len(strats)
    _STATE = WorkerState(strats, config, game_profile, table, ids, device)


_TABLE_CACHE: Dict[tuple, tuple] = {}


def _prepared_table(strategies: Sequence[ThresholdStrategy]):
    """Id resolution + packing of a strategy list, memoised on its contents: a runner calls
    ``_init_worker`` once per (root, k) cell with the same 5,160-strategy grid."""
    key = tuple((s.score_threshold, s.dice_threshold, s.smart_five, s.smart_one, s.consider_score,
                 s.consider_dice, s.require_both, s.auto_hot_dice, s.run_up_score,
                 s.favor_dice_or_score, s.strategy_id) for s in strategies)
    hit = _TABLE_CACHE.get(key)
    if hit is None:
        strats = _prepare_public_helper_strategies(strategies)
        hit = (strats, pack_strategies(strats), np.array([s.strategy_id for s in strats], dtype=np.int64))
        if len(_TABLE_CACHE) >= 2:
            _TABLE_CACHE.clear()
        _TABLE_CACHE[key] = hit
    return hit


def _coerce_shuffle_task(task: ShuffleTask | int) -> ShuffleTask:
    if isinstance(task, ShuffleTask):
        return task
    k = _STATE.cfg.n_players if _STATE is not None else 0
    return ShuffleTask(root_seed=int(task), k=k, shuffle_index=0, shuffle_seed=int(task),
                       deterministic_batch_id=0)


def _require_state() -> WorkerState:
    if _STATE is None:
        raise RuntimeError("_init_worker has not been called")
    return _STATE


def _contiguous_runs(tasks: Sequence[ShuffleTask]) -> Iterator[Tuple[int, int]]:
    """``(start, stop)`` slices of ``tasks`` that are one (root, k) with consecutive shuffles."""
    start = 0
    for i in range(1, len(tasks) + 1):
        if (i == len(tasks) or tasks[i].root_seed != tasks[start].root_seed
                or tasks[i].k != tasks[start].k
                or tasks[i].shuffle_index != tasks[i - 1].shuffle_index + 1):
            yield start, i
            start = i


def _launch_run(state: WorkerState, tasks: Sequence[ShuffleTask], *, want_rows: bool):
    """One ``fb_play_tournament`` launch for a contiguous shuffle run.

    Returns ``(tallies[n, 26], (attempted, completed, safety_limit), rows | None,
    first_seen[n, 4])`` as host values for the whole run.
    """
    from .device import get_engine

    first = tasks[0]
    k, n = first.k, len(tasks)
    if k != state.cfg.n_players:
        raise ValueError(f"task k={k} does not match the worker's n_players={state.cfg.n_players}")
    prof = state.game_profile
    eng = get_engine(state.device)
    res = eng.play_tournament(
        first.root_seed, k, first.shuffle_index, n, state.table,
        target_score=prof.default_target_score if prof else 10_000,
        max_rounds=prof.default_max_rounds if prof else 200,
        overrides=(prof.tournament_overrides_for(first.root_seed, k, first.shuffle_index, n)
                   if prof else ()),
        want_rows=want_rows, want_game_seeds=want_rows, want_first_seen=True)
    tallies = res.tallies.cpu().numpy()[0]
    first_seen = res.first_seen.cpu().numpy()
    totals = res.totals.cpu().numpy()
    rows = res.rows_numpy() if res.rows is not None else None
    if rows is not None:
        check_row_flags(rows)
    elif totals[7]:
        raise RuntimeError("a game hit ROLL_LIMIT or overflowed an int16 row counter")
    return tallies, (int(totals[0]), int(totals[1]), int(totals[2])), rows, first_seen


def _expand_shuffle_rows(state: WorkerState, task: ShuffleTask, rows: np.ndarray) -> List[Dict[str, Any]]:
    """Row mappings of one shuffle with the reference's provenance block (:350-361)."""
    rows = rows.copy()
    rows["seats"]["strategy"] = state.ids[rows["seats"]["strategy"]]
    prov = [{
        "root_seed": task.root_seed, "k": task.k, "shuffle_index": task.shuffle_index,
        "game_index": g, "deterministic_batch_id": task.deterministic_batch_id,
        "shuffle_seed": task.shuffle_seed, "game_seed": int(rows["game_seed"][g]),
        "rng_scheme_version": RNG_SCHEME_VERSION,
        "rng_purpose_namespace": int(RandomPurpose.TOURNAMENT_GAME),
    } for g in range(len(rows))]
    return expand_rows(rows, prov)


def _play_one_shuffle(task: ShuffleTask | int, *, collect_rows: bool = False) -> Tuple[
        Counter, MetricSums, MetricSums, List[Dict[str, Any]]]:
    """Play all games of one shuffle and aggregate the results (run_tournament.py:301-393)."""
    state = _require_state()
    work = _coerce_shuffle_task(task)
    tallies, games, rows, seen = _launch_run(state, [work], want_rows=collect_rows)
    wins, sums, sq_sums = tallies_to_outcome(tallies, state.ids, games, seen)
    out_rows = _expand_shuffle_rows(state, work, rows) if collect_rows else []
    return wins, sums, sq_sums, out_rows


def _play_shuffle(task: ShuffleTask | int) -> Counter:
    wins, _, _, _ = _play_one_shuffle(task, collect_rows=False)
    return wins


def _run_chunk(shuffle_tasks: Sequence[ShuffleTask | int]) -> Counter:
    """Play a batch of shuffles and tally wins (run_tournament.py:403-457): one launch per
    contiguous run instead of one Python game loop per shuffle."""
    state = _require_state()
    tasks = [_coerce_shuffle_task(t) for t in shuffle_tasks]
    total = OutcomeCounter()
    for a, b in _contiguous_runs(tasks):
        tallies, games, _, seen = _launch_run(state, tasks[a:b], want_rows=False)
        wins, _, _ = tallies_to_outcome(tallies, state.ids, games, seen)
        total.absorb(wins)
    return total


def _run_chunk_item(item: Tuple[int, Sequence[ShuffleTask]], *,
                    chunk_fn: Callable[[Sequence[ShuffleTask]], object]
                    ) -> Tuple[int, tuple, object]:
    chunk_index, seeds = item
    tasks = tuple(seeds)
    return chunk_index, tasks, chunk_fn(tasks)


def _atomic_write(path: Path, write: Callable[[Path], None]) -> None:
    """temp -> fsync -> rename in the target directory (utils/writer.py:41-124 semantics)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    fd, tmp = tempfile.mkstemp(prefix=f".{path.name}.", suffix=".tmp", dir=path.parent)
    os.close(fd)
    try:
        write(Path(tmp))
        with open(tmp, "rb") as fh:
            os.fsync(fh.fileno())
        os.replace(tmp, path)
    finally:
        if os.path.exists(tmp):
            os.unlink(tmp)


def _append_manifest(manifest_path: Path, record: Mapping[str, Any]) -> None:
    """One NDJSON line per shard (utils/manifest.py:134-166)."""
    _append_manifest_lines(manifest_path, [record])


def _append_manifest_lines(manifest_path: Path, records: Sequence[Mapping[str, Any]]) -> None:
    """Several manifest lines with one open / flush / fsync: every line still names a shard that is
    already on disk (the records of a finished writer task), but 4,300 fsyncs of the manifest per
    cell were the slowest part of rows mode after the Parquet encode."""
    if not records:
        return
    manifest_path.parent.mkdir(parents=True, exist_ok=True)
    with open(manifest_path, "a", encoding="utf-8") as fh:
        fh.write("".join(json.dumps(rec, sort_keys=True, separators=(",", ":")) + "\n" for rec in records))
        fh.flush()
        os.fsync(fh.fileno())


def shard_manifest_extra(state: WorkerState, task: ShuffleTask, shard_name: str) -> Dict[str, Any]:
    """The manifest fields the reference records per row shard (run_tournament.py:536-557)."""
    record: Dict[str, Any] = {
        "path": shard_name, "root_seed": task.root_seed, "n_players": task.k,
        "shuffle_index": task.shuffle_index, "shuffle_seed": task.shuffle_seed,
        "deterministic_batch_id": task.deterministic_batch_id,
        "rng_scheme_version": RNG_SCHEME_VERSION,
        "rng_purpose_namespace": int(RandomPurpose.TOURNAMENT_SHUFFLE),
        "outcome_schema_version": OUTCOME_SCHEMA_VERSION,
        "tournament_method_version": TOURNAMENT_METHOD_VERSION, "pid": os.getpid(),
    }
    if state.game_profile is not None:
        record["game_profile_sha256"] = state.game_profile.sha256
    return record


# Row shards are small (one shuffle: 2,580 rows x 46 columns for k=2 on the full grid) and the host
# encode is what bounds rows mode end to end (bench.py `e2e_parquet`).  pyarrow's default builds a
# dictionary for every column and falls back to plain pages when it overflows, which costs more
# than it saves on the per-seat counters and scores: dictionaries only for the columns that are
# constant or nearly so inside a shard.  Same schema, same values, snappy like the reference's writer
# (utils/writer.py:46); 42 % less encode time per shard, files 19 % larger.
_LOW_CARDINALITY = ("root_seed", "k", "shuffle_index", "deterministic_batch_id", "shuffle_seed",
                    "termination_status", "hit_safety_limit", "outcome_schema_version", "winner_seat",
                    "rng_scheme_version", "rng_purpose_namespace", "seat_ranks")


def row_shard_write_options(table) -> Dict[str, Any]:
    """``pq.write_table`` keyword arguments for a row shard (see above)."""
    cols = [n for n in table.schema.names
            if n in _LOW_CARDINALITY or n.endswith("_rank") or n.endswith("_hit_max_rounds")]
    return {"use_dictionary": cols, "compression": "snappy"}


def _write_shard(out: Path, manifest_file: Path, table, manifest_extra: Mapping[str, Any]) -> None:
    """Default shard writer: Parquet temp->fsync->rename, then one manifest line."""
    import pyarrow.parquet as pq

    opts = row_shard_write_options(table)
    _atomic_write(out, lambda p: pq.write_table(table, p, **opts))
    _append_manifest(manifest_file, {"path": out.name, "rows": table.num_rows, **manifest_extra})


# Where rows mode spends its host time (reset by the caller; bench.py's e2e_parquet leg reads it):
# wall seconds inside the launches (kernels + D2H of the rows), CPU-seconds of the Arrow build and
# of the Parquet encode + fsync + rename summed over the writer threads.
IO_STATS: Dict[str, float] = {"launch_wall_s": 0.0, "arrow_build_cpu_s": 0.0, "parquet_write_cpu_s": 0.0,
                              "shards": 0}
_IO_LOCK = __import__("threading").Lock()


def _now() -> float:
    import time

    return time.perf_counter()


def _run_chunk_metrics(shuffle_tasks: Sequence[ShuffleTask | int], *, collect_rows: bool = False,
                       row_dir: Path | None = None, manifest_path: Path | None = None,
                       row_sidecar: object | None = None,
                       shard_writer: Callable[[Path, Path, Any, Mapping[str, Any]], None] | None = None,
                       io_threads: int | None = None) -> Tuple[Counter, MetricSums, MetricSums]:
    """Play shuffles and accumulate metrics (run_tournament.py:473-585).

    In rows mode every shuffle leaves one Parquet shard
    ``rows_{root}_{k}p_{shuffle:012d}.parquet`` plus one manifest line before returning.
    ``shard_writer(out_path, manifest_path, arrow_table, manifest_extra)`` publishes a shard; the
    default writes temp->rename without a sidecar, ``reference_shim`` passes the reference's own
    ``run_streaming_shard`` so that hash-bound sidecars come out exactly as the reference
    writes them.  ``row_sidecar`` without such a writer is refused: the artifact contract is
    outside this path.

    With the default writer the Arrow build and the Parquet encode of the shards run on
    ``io_threads`` host threads (default: the host cores, at most 32; pyarrow and numpy release the
    GIL) while the next launch is already playing; manifest lines are appended in shuffle order once
    their shard is on disk.  A custom ``shard_writer`` is called sequentially, in order.
    """
    if row_sidecar is not None and shard_writer is None:
        raise NotImplementedError("row_sidecar needs a shard_writer that implements the artifact contract")
    state = _require_state()
    tasks = [_coerce_shuffle_task(t) for t in shuffle_tasks]
    wins_total = OutcomeCounter()
    sums_total: MetricSums = {m: defaultdict(float) for m in METRIC_LABELS}
    sq_total: MetricSums = {m: defaultdict(float) for m in METRIC_LABELS}
    want_rows = collect_rows and row_dir is not None
    gps = state.cfg.games_per_shuffle
    manifest_file = Path(manifest_path or (Path(row_dir) / "manifest.jsonl")) if want_rows else None
    threads = io_threads if io_threads is not None else min(32, os.cpu_count() or 1)
    pool = None
    pending: List[Tuple[Any, Sequence[Path], Sequence[Mapping[str, Any]]]] = []  # (future, shard paths, manifest records)

    def shard_table(shard: np.ndarray, task: ShuffleTask):
        return compact_rows_to_table(
            shard, root_seed=task.root_seed, k=task.k, shuffle_index=task.shuffle_index,
            game_index=np.arange(gps), deterministic_batch_id=task.deterministic_batch_id,
            shuffle_seed=task.shuffle_seed)

    def build_and_write(block: np.ndarray, group: Sequence[ShuffleTask], outs: Sequence[Path]) -> List[int]:
        """One writer task = a run of consecutive shuffles (up to a deterministic batch): ONE Arrow
        build for all their rows (numpy and Arrow release the GIL on arrays this size; per-shuffle
        builds spend their time in interpreter overhead and serialise on it), then one zero-copy
        slice and one Parquet file per shuffle."""
        import time

        import pyarrow.parquet as pq

        t0 = time.perf_counter()
        per = np.repeat(np.arange(len(group)), gps)
        tbl = compact_rows_to_table(
            block, root_seed=group[0].root_seed, k=group[0].k,
            shuffle_index=np.array([t.shuffle_index for t in group], dtype=np.int64)[per],
            game_index=np.tile(np.arange(gps, dtype=np.int32), len(group)),
            deterministic_batch_id=np.array([t.deterministic_batch_id for t in group], dtype=np.int32)[per],
            shuffle_seed=np.array([t.shuffle_seed for t in group], dtype=np.int64)[per])
        t1 = time.perf_counter()
        counts = []
        opts = row_shard_write_options(tbl)
        for i, out in enumerate(outs):
            part = tbl.slice(i * gps, gps)
            _atomic_write(out, lambda p, part=part: pq.write_table(part, p, **opts))
            counts.append(part.num_rows)
        t2 = time.perf_counter()
        with _IO_LOCK:
            IO_STATS["arrow_build_cpu_s"] += t1 - t0
            IO_STATS["parquet_write_cpu_s"] += t2 - t1
            IO_STATS["shards"] += len(outs)
        return counts

    def drain() -> None:
        for fut, outs, extras in pending:
            _append_manifest_lines(manifest_file, [{"path": out.name, "rows": n_rows, **extra}
                                                   for out, extra, n_rows in zip(outs, extras, fut.result())])
        pending.clear()

    try:
        for a, b in _contiguous_runs(tasks):
            run = tasks[a:b]
            t_launch = _now()
            tallies, games, rows, seen = _launch_run(state, run, want_rows=want_rows)
            IO_STATS["launch_wall_s"] += _now() - t_launch
            wins, sums, sqs = tallies_to_outcome(tallies, state.ids, games, seen)
            wins_total.absorb(wins)
            _add_sums(sums_total, sums)
            _add_sums(sq_total, sqs)
            if not want_rows:
                continue
            rows = rows.copy()
            rows["seats"]["strategy"] = state.ids[rows["seats"]["strategy"]]
            drain()                 # the previous launch's shards (written while this one played)
            outs = [Path(row_dir) / f"rows_{t.root_seed}_{t.k}p_{t.shuffle_index:012d}.parquet" for t in run]
            extras = [shard_manifest_extra(state, t, o.name) for t, o in zip(run, outs)]
            if shard_writer is not None:
                for i, task in enumerate(run):
                    shard_writer(outs[i], manifest_file, shard_table(rows[i * gps:(i + 1) * gps], task), extras[i])
                continue
            # writer tasks: runs of consecutive shuffles of one deterministic batch, at most 64 each
            cuts = [0]
            for i in range(1, len(run)):
                if run[i].deterministic_batch_id != run[cuts[-1]].deterministic_batch_id or i - cuts[-1] >= 64:
                    cuts.append(i)
            cuts.append(len(run))
            if pool is None and threads > 1:
                from concurrent.futures import ThreadPoolExecutor

                pool = ThreadPoolExecutor(max_workers=threads, thread_name_prefix="fb-shard")
            for i0, i1 in zip(cuts[:-1], cuts[1:]):
                args = (rows[i0 * gps:i1 * gps], run[i0:i1], outs[i0:i1])
                if pool is None:
                    _append_manifest_lines(manifest_file, [
                        {"path": out.name, "rows": n_rows, **extra}
                        for out, extra, n_rows in zip(outs[i0:i1], extras[i0:i1], build_and_write(*args))])
                else:
                    pending.append((pool.submit(build_and_write, *args), outs[i0:i1], extras[i0:i1]))
        drain()
    finally:
        if pool is not None:
            pool.shutdown(wait=True)
    return wins_total, sums_total, sq_total


# --------------------------------------------------------------------------- cell runner
def make_shuffle_tasks(root_seed: int, k: int, shuffle_indices: Sequence[int],
                       deterministic_batch_size: int) -> List[ShuffleTask]:
    """Tasks with the purpose-100 fingerprint, as ``_iter_original_chunk_items`` builds them
    (run_tournament.py:944-984)."""
    return [ShuffleTask(root_seed=root_seed, k=k, shuffle_index=s,
                        shuffle_seed=coordinate_seed(RandomPurpose.TOURNAMENT_SHUFFLE,
                                                     root_seed=root_seed, k=k, shuffle_index=s,
                                                     dtype=np.uint32),
                        deterministic_batch_id=s // deterministic_batch_size)
            for s in shuffle_indices]


def shard_batches(num_shuffles: int, batch_size: int, rank: int, world: int
                  ) -> List[Tuple[int, int, int]]:
    """``(batch_id, shuffle0, n_shuffles)`` of the deterministic batches rank ``rank`` owns.

    Batch b is shuffles [b*B, min((b+1)*B, S)), the reference's recovery unit
    (run_tournament.py:974).  The batches are dealt in contiguous blocks -- the first
    ``n_batches % world`` ranks take one more -- so that a rank's share of a cell is ONE launch
    (``merge_ranges``) large enough to fill the GPU, not one launch per batch."""
    if batch_size < 1 or world < 1 or not 0 <= rank < world:
        raise ValueError("bad batch size / rank / world")
    n_batches = -(-num_shuffles // batch_size)
    base, extra = divmod(n_batches, world)
    first = rank * base + min(rank, extra)
    return [(b, b * batch_size, min(batch_size, num_shuffles - b * batch_size))
            for b in range(first, first + base + (1 if rank < extra else 0))]


def merge_ranges(batches: Sequence[Tuple[int, int, int]]) -> List[Tuple[int, int]]:
    """Coalesce owned batches into maximal contiguous ``(shuffle0, n_shuffles)`` launches."""
    out: List[Tuple[int, int]] = []
    for _, s0, n in batches:
        if out and out[-1][0] + out[-1][1] == s0:
            out[-1] = (out[-1][0], out[-1][1] + n)
        else:
            out.append((s0, n))
    return out


def all_reduce_tallies(tallies, totals):
    """Sum the small tally / totals tensors over ranks (the path's ONE exchange step)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tallies)
        dist.all_reduce(totals)
    return tallies, totals


def run_cell(root_seed: int, k: int, num_shuffles: int, table, *, batch_size: int,
             launch: Callable[..., Tuple[Any, ...]], rank: int = 0, world: int = 1,
             max_shuffles_per_launch: int = 1 << 20, want_first_seen: bool = False):
    """Play this rank's share of a (root, k) cell and merge.

    ``launch(root_seed, k, shuffle0, n_shuffles, table, tallies, totals)`` accumulates into
    the ``int64`` tensors it is given (``Engine.play_tournament`` on a GPU; the tests pass a
    CPU stand-in).  Returns the merged ``(tallies[1, n, 26], totals[20])`` tensors; with
    ``want_first_seen`` the launch also returns its ``first_seen`` ordinals and the result
    gains the cell-wide ``int64 [n, 4]`` first-seen table (MIN over launches and ranks; the
    key order of the reference's counters, see ``tallies_to_outcome``).
    """
    import torch
    import torch.distributed as dist

    from .layout import TOTALS_WIDTH

    n = len(table) if isinstance(table, np.ndarray) else table.numel() // 8
    never = torch.iinfo(torch.int64).max
    tallies = totals = seen = None
    for s0, cnt in merge_ranges(shard_batches(num_shuffles, batch_size, rank, world)):
        while cnt > 0:
            step = min(cnt, max_shuffles_per_launch)
            out = launch(root_seed, k, s0, step, table, tallies, totals)
            tallies, totals = out[0], out[1]
            if want_first_seen:
                local = out[2].to(torch.int64)          # int32, -1 = never, else ordinal in the launch
                glob = torch.where(local < 0, torch.full_like(local, never),
                                   (local & 0xFFFFFFFF) + s0 * n)
                seen = glob if seen is None else torch.minimum(seen, glob)
            s0 += step
            cnt -= step
    dev = table.device if hasattr(table, "device") else "cpu"
    if tallies is None:  # a rank that owns no batch still takes part in the reduction
        tallies = torch.zeros((1, n, TALLY_WIDTH), dtype=torch.int64, device=dev)
        totals = torch.zeros(TOTALS_WIDTH, dtype=torch.int64, device=dev)
    tallies, totals = all_reduce_tallies(tallies, totals)
    if not want_first_seen:
        return tallies, totals
    if seen is None:
        seen = torch.full((n, 4), never, dtype=torch.int64, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(seen, op=dist.ReduceOp.MIN)
    return tallies, totals, seen


# ---------------------------------------------------------------------------------------------
# Several cells over several GPUs: strong scaling of a fixed workload (configs/farkle_mega_config.yaml:
# k in {2,3,4,5,6,8,10,12}; SURVEY.md section 8e "balance by estimated rolls")
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CellSegment:
    """A rank's contiguous share of one cell: whole deterministic batches."""

    cell: int          # index into the cell list
    root_seed: int
    k: int
    shuffle0: int
    n_shuffles: int


# Measured on one B200 (profiles/r02_cells.log + the 1.4-1.7 ms of permutation, seeding, finish and
# tally passes per cell, profiles/r02_timeline.md): a full-grid 4,300-shuffle cell takes
# 13.6 + 12.4 / k ms within 3 % for k = 2, 4, 6, 12 (19.6, 17.3, 15.5, 14.4): seat-exposures are
# constant per cell, a seat's turns cost the same at every table size, and the per-game part shrinks
# with the game count.  (Round 1's kernels: 13.9 + 18.2 / k.)
CELL_MS_CONST, CELL_MS_PER_K = 13.6, 12.4
CELL_MS_REFERENCE_EXPOSURES = 5160 * 4300
SEGMENT_OVERHEAD_MS = 0.5       # kernel start-up and drain tail of one more launch


def cell_cost_ms(k: int, n_strategies: int, n_shuffles: int) -> float:
    """Estimated single-GPU time of ``n_shuffles`` shuffles of a k-player cell."""
    scale = n_strategies * n_shuffles / CELL_MS_REFERENCE_EXPOSURES
    return (CELL_MS_CONST + CELL_MS_PER_K / k) * scale


def plan_cells(cells: Sequence[Tuple[int, int, int]], n_strategies: int, world: int, *, batch_size: int,
               cost: Callable[[int, int, int], float] = cell_cost_ms,
               segment_overhead_ms: float = SEGMENT_OVERHEAD_MS) -> List[List[CellSegment]]:
    """Deal the cells ``(root_seed, k, num_shuffles)`` of a run to ``world`` ranks.

    The unit of ownership stays the deterministic batch (run_tournament.py:974), so every rank plays
    whole batches and disjoint coordinate ranges.  Cells are laid end to end, most expensive first,
    and the line is cut into ``world`` pieces of (nearly) equal estimated time: a rank gets whole
    cells plus at most two partial ones, instead of 1/world of EVERY cell (which leaves each launch
    too small to fill a GPU: 4.4x on 8 GPUs in round 1).  The smallest makespan for which the greedy
    fill fits is found by bisection; every extra launch is charged ``segment_overhead_ms``.
    Deterministic: every rank computes the same plan.
    """
    if world < 1 or batch_size < 1:
        raise ValueError("bad world / batch size")
    order = sorted(range(len(cells)), key=lambda i: (-cost(cells[i][1], n_strategies, cells[i][2]), i))
    per_batch = {i: cost(cells[i][1], n_strategies, min(batch_size, cells[i][2]) or 1) for i in order}
    n_batches = {i: -(-cells[i][2] // batch_size) for i in order}

    def fill(limit: float) -> List[List[CellSegment]] | None:
        plan: List[List[CellSegment]] = [[] for _ in range(world)]
        rank, room = 0, limit
        for i in order:
            root, k, shuffles = cells[i]
            done = 0
            while done < n_batches[i]:
                take = min(n_batches[i] - done, int((room - segment_overhead_ms) / per_batch[i] + 1e-9))
                if take < 1:
                    rank += 1
                    room = limit
                    if rank == world:
                        return None
                    continue
                s0 = done * batch_size
                plan[rank].append(CellSegment(i, root, k, s0, min(take * batch_size, shuffles - s0)))
                room -= segment_overhead_ms + take * per_batch[i]
                done += take
        return plan

    total = sum(per_batch[i] * n_batches[i] for i in order)
    lo, hi = total / world, total + (len(cells) + world) * segment_overhead_ms + 1.0
    best = fill(hi)
    assert best is not None
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        got = fill(mid)
        if got is None:
            lo = mid
        else:
            best, hi = got, mid
    return best


def run_cells(cells: Sequence[Tuple[int, int, int]], table, *, batch_size: int, play_cells: Callable[..., None],
              rank: int = 0, world: int = 1, plan: Sequence[Sequence[CellSegment]] | None = None):
    """Play this rank's share of ALL cells, then merge with ONE all-reduce.

    ``play_cells(segments, table)`` receives ``(root_seed, k, shuffle0, n_shuffles, tallies, totals)``
    tuples (``Engine.play_cells`` on a GPU: pipelined launches; the CPU tests pass a stand-in) and
    accumulates into the int64 views it is given.  Returns ``(tallies [cells, 1, n, 26], totals
    [cells, 20])`` summed over ranks: both live in one flat buffer, so the path's exchange step is
    a single collective for the whole run instead of one per cell.
    """
    import torch
    import torch.distributed as dist

    from .layout import TOTALS_WIDTH

    n = len(table) if isinstance(table, np.ndarray) else table.numel() // 8
    dev = table.device if hasattr(table, "device") else "cpu"
    n_cells = len(cells)
    flat = torch.zeros(n_cells * (n * TALLY_WIDTH + TOTALS_WIDTH), dtype=torch.int64, device=dev)
    tallies = flat[: n_cells * n * TALLY_WIDTH].view(n_cells, 1, n, TALLY_WIDTH)
    totals = flat[n_cells * n * TALLY_WIDTH:].view(n_cells, TOTALS_WIDTH)
    if plan is None:
        plan = plan_cells(cells, n, world, batch_size=batch_size)
    mine = plan[rank]
    if mine:
        play_cells([(sg.root_seed, sg.k, sg.shuffle0, sg.n_shuffles, tallies[sg.cell], totals[sg.cell])
                    for sg in mine], table)
    # `world` (not the size of the default process group) decides: a caller that plays a plan alone
    # (world == 1) inside a multi-rank job must not enter a collective the other ranks do not join
    if world > 1 and dist.is_available() and dist.is_initialized():
        dist.all_reduce(flat)
    return tallies, totals


def _engine_launch(eng, want_kw: Mapping[str, Any]):
    def launch(root_seed, k, shuffle0, n_shuffles, table, tallies, totals):
        res = eng.play_tournament(root_seed, k, shuffle0, n_shuffles, table, tallies=tallies,
                                  totals=totals, want_first_seen=True, **want_kw)
        return res.tallies, res.totals, res.first_seen
    return launch


ROWS_LAUNCH_BATCHES = 8      # deterministic batches per launch in rows mode (fills the GPU)


def _checkpoint_meta(cfg: TournamentConfig, global_seed: int, profile: GameProfile | None,
                     extra: Mapping[str, Any] | None) -> Dict[str, Any]:
    """The reference's checkpoint ``meta`` block, key for key (run_tournament.py:1139-1171)."""
    meta: Dict[str, Any] = {
        "n_players": cfg.n_players, "num_shuffles": cfg.num_shuffles, "global_seed": global_seed,
        "n_strategies": cfg.n_strategies, "rng_scheme_version": RNG_SCHEME_VERSION,
        "outcome_schema_version": OUTCOME_SCHEMA_VERSION,
        "tournament_method_version": TOURNAMENT_METHOD_VERSION, "rng_bit_generator": "PCG64DXSM",
        "coordinate_contract_version": 2,
        "shuffle_purpose_namespace": int(RandomPurpose.TOURNAMENT_SHUFFLE),
        "shuffle_permutation_purpose_namespace": int(RandomPurpose.SHUFFLE_PERMUTATION),
        "game_purpose_namespace": int(RandomPurpose.TOURNAMENT_GAME),
        "player_purpose_namespace": int(RandomPurpose.TOURNAMENT_PLAYER),
        "deterministic_batch_size": cfg.deterministic_batch_size,
    }
    if profile is not None:
        meta["game_profile_sha256"] = profile.sha256
    if extra:
        meta.update(extra)
    return meta


def _save_checkpoint(path: Path, wins: OutcomeCounter, sums: MetricSums | None, sqs: MetricSums | None,
                     meta: Mapping[str, Any]) -> None:
    """run_tournament.py:622-651 (without the sidecar: the artifact contract is outside this path)."""
    payload: Dict[str, Any] = {"win_totals": wins, "outcome_counts": wins.outcome_payload()}
    if sums is not None and sqs is not None:
        payload["metric_sums"] = {m: dict(v) for m, v in sums.items()}
        payload["metric_square_sums"] = {m: dict(v) for m, v in sqs.items()}
    if meta:
        payload["meta"] = dict(meta)
    _atomic_write(path, lambda p: p.write_bytes(pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL)))


def _load_checkpoint(path: Path) -> Tuple[OutcomeCounter, MetricSums | None, MetricSums | None, set[int]]:
    """Aggregates and completed shuffles of an earlier run (run_tournament.py:1243-1330)."""
    payload = pickle.loads(path.read_bytes())
    meta = payload.get("meta", {}) if isinstance(payload, Mapping) else {}
    if "rng_scheme_version" in meta and meta["rng_scheme_version"] != RNG_SCHEME_VERSION:
        raise ValueError("Checkpoint RNG scheme is stale or unsupported; restart from a v2 output root")
    wins = payload["win_totals"]
    if not isinstance(wins, OutcomeCounter):
        wins = _restore_outcome_counter(dict(wins), payload.get("outcome_counts") or {})

    def coerce(raw: Any) -> MetricSums | None:
        if raw is None:
            return None
        return {m: defaultdict(float, raw.get(m, {})) for m in METRIC_LABELS}

    return (wins, coerce(payload.get("metric_sums")), coerce(payload.get("metric_square_sums")),
            {int(v) for v in meta.get("completed_shuffle_indices", [])})


def _add_sums(total: MetricSums, part: MetricSums) -> None:
    for label in METRIC_LABELS:
        for key, v in part[label].items():
            total[label][key] += v


def _write_metric_chunk(directory: Path, state: WorkerState, root: int, tasks: Sequence[ShuffleTask],
                        wins: OutcomeCounter, sums: MetricSums, sqs: MetricSums) -> None:
    """One ``metrics_{chunk:06d}.parquet`` + manifest line per deterministic batch
    (run_tournament.py:1603-1668; chunk indices are 1-based, :944-984)."""
    import pyarrow as pa

    chunk_index = tasks[0].deterministic_batch_id + 1
    rows = []
    for label in METRIC_LABELS:
        for strat in sorted(set(sums[label]) | set(wins.attempted_exposures), key=str):
            rows.append({"metric": label, "strategy": strat, "sum": sums[label].get(strat, 0.0),
                         "square_sum": sqs[label][strat] if strat in sqs[label] else 0.0,
                         "wins": int(wins.get(strat, 0)),
                         "attempted_exposures": int(wins.attempted_exposures.get(strat, 0)),
                         "completed_exposures": int(wins.completed_exposures.get(strat, 0)),
                         "safety_limit_exposures": int(wins.safety_limit_exposures.get(strat, 0))})
    tbl = pa.Table.from_pylist(rows)
    out = directory / f"metrics_{chunk_index:06d}.parquet"
    extra: Dict[str, Any] = {
        "chunk_index": chunk_index, "process_block_index": chunk_index, "root_seed": root,
        "n_players": state.cfg.n_players, "deterministic_batch_id": tasks[0].deterministic_batch_id,
        "shuffle_index_start": tasks[0].shuffle_index, "shuffle_index_end": tasks[-1].shuffle_index,
        "shuffle_count": len(tasks), "shuffle_indices": [t.shuffle_index for t in tasks],
        "shuffle_seeds": [t.shuffle_seed for t in tasks], "rng_scheme_version": RNG_SCHEME_VERSION,
        "rng_purpose_namespace": int(RandomPurpose.TOURNAMENT_SHUFFLE),
        "outcome_schema_version": OUTCOME_SCHEMA_VERSION,
        "tournament_method_version": TOURNAMENT_METHOD_VERSION}
    if state.game_profile is not None:
        extra["game_profile_sha256"] = state.game_profile.sha256
    _write_shard(out, directory / "metrics_manifest.jsonl", tbl, extra)


def _load_metric_chunk_aggregates(manifest_path: Path, k: int) -> Tuple[OutcomeCounter, MetricSums, MetricSums]:
    """Aggregates rebuilt from the metric-chunk shards a manifest lists, in manifest order
    (run_tournament.py:866-941).  This is what the reference's final checkpoint holds when a run
    with a metric-chunk directory started without checkpointed sums: every strategy that was
    seated appears, zeros included, keyed in the chunk tables' order."""
    import pyarrow.parquet as pq

    wins = OutcomeCounter()
    sums: MetricSums = {m: defaultdict(float) for m in METRIC_LABELS}
    sqs: MetricSums = {m: defaultdict(float) for m in METRIC_LABELS}
    for line in manifest_path.read_text().splitlines():
        if not line.strip():
            continue
        chunk = manifest_path.parent / json.loads(line)["path"]
        if not chunk.exists():
            raise FileNotFoundError(f"Missing metric chunk listed in manifest: {chunk}")
        for row in pq.read_table(chunk).to_pylist():
            label, strategy = row["metric"], row["strategy"]
            sums[label][strategy] += float(row["sum"])
            sqs[label][strategy] += float(row["square_sum"])
            if label == METRIC_LABELS[0]:
                wins[strategy] += int(row["wins"])
                wins.attempted_exposures[strategy] += int(row["attempted_exposures"])
                wins.completed_exposures[strategy] += int(row["completed_exposures"])
                wins.safety_limit_exposures[strategy] += int(row["safety_limit_exposures"])
    wins.games_completed = int(sum(wins.values()))
    wins.games_attempted = int(sum(wins.attempted_exposures.values())) // k
    wins.games_safety_limit = int(sum(wins.safety_limit_exposures.values())) // k
    return wins, sums, sqs


def _manifest_int_set(path: Path, key: str) -> set[int]:
    out: set[int] = set()
    if path.exists():
        for line in path.read_text().splitlines():
            if line.strip():
                rec = json.loads(line)
                if key in rec:
                    out.add(int(rec[key]))
    return out


def run_tournament(*, config: TournamentConfig | None = None, global_seed: int = 0,
                   checkpoint_path: Path | str = "checkpoint.pkl", n_jobs: int | None = None,
                   collect_metrics: bool = False, row_output_directory: Path | None = None,
                   metric_chunk_directory: Path | None = None,
                   num_shuffles: int = NUM_SHUFFLES,
                   strategies: Sequence[ThresholdStrategy] | None = None, resume: bool = True,
                   checkpoint_metadata: Mapping[str, Any] | None = None,
                   oracle_game_profile: GameProfile | None = None,
                   write_final_metrics_artifact: bool = True, device: int | None = None,
                   io_threads: int | None = None) -> None:
    """Run one (root, k) tournament cell on the GPU(s) (run_tournament.py:1050-1859).

    Keeps the reference's observable results: the checkpoint pickle ``{"win_totals":
    OutcomeCounter, "outcome_counts", "metric_sums", "metric_square_sums", "meta"}`` (``meta`` key
    for key, with ``completed_shuffle_indices`` / ``completed_process_block_indices``),
    ``{k}p_metrics.parquet``, with ``row_output_directory`` one Parquet shard + manifest line per
    shuffle, with ``metric_chunk_directory`` one ``metrics_{chunk}.parquet`` + manifest line per
    deterministic batch, a checkpoint every ``config.ckpt_every_sec`` seconds, and resume: shuffles
    listed in the checkpoint are not replayed, shuffles whose row shard is already in the manifest
    are not rewritten (their tallies are replayed on the GPU, which costs microseconds, instead of
    re-reading the shards as the reference does, :820-863).  Every game is played ONCE: rows and
    tallies come out of the same launches.  ``n_jobs`` is accepted for compatibility: parallelism
    is the GPU's (and ``io_threads`` Parquet writers on the host).

    Two routes.  Plain tallies (no rows, no metric chunks, nothing to resume): one launch per rank
    for the whole cell, one all-reduce of the tally tensor.  Otherwise the cell is walked in groups
    of deterministic batches; ranks own contiguous blocks of batches and their partial counters are
    joined in rank (= batch) order.
    """
    import time

    import torch
    import torch.distributed as dist

    from .device import get_engine

    del n_jobs
    if strategies is None:
        strategies, _ = generate_strategy_grid()
    cfg = config or TournamentConfig()
    if num_shuffles != cfg.num_shuffles:
        cfg.num_shuffles = num_shuffles
    if cfg.deterministic_batch_size < 1:
        raise ValueError("deterministic_batch_size must be positive")
    _init_worker(strategies, cfg, oracle_game_profile, device=device)
    state = _require_state()
    k, root = cfg.n_players, int(global_seed)
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    ckpt_path = Path(checkpoint_path)
    collect_rows = row_output_directory is not None
    want_sums = collect_metrics or collect_rows
    meta = _checkpoint_meta(cfg, root, oracle_game_profile, checkpoint_metadata)
    batch = cfg.deterministic_batch_size
    n_batches = -(-cfg.num_shuffles // batch)

    wins_prev, sums_prev, sqs_prev, completed = OutcomeCounter(), None, None, set()
    if resume and ckpt_path.exists():
        wins_prev, sums_prev, sqs_prev, completed = _load_checkpoint(ckpt_path)
        completed &= set(range(cfg.num_shuffles))
    row_dir = Path(row_output_directory) if collect_rows else None
    row_manifest = row_dir / "manifest.jsonl" if row_dir else None
    rows_done = _manifest_int_set(row_manifest, "shuffle_index") if (row_manifest and resume) else set()
    chunk_dir = Path(metric_chunk_directory) if metric_chunk_directory is not None else None
    chunks_done = (_manifest_int_set(chunk_dir / "metrics_manifest.jsonl", "chunk_index")
                   if (chunk_dir and resume) else set())
    if chunk_dir:
        chunk_dir.mkdir(parents=True, exist_ok=True)

    plain = not collect_rows and chunk_dir is None and not completed
    if plain:
        eng = get_engine(state.device)
        prof = state.game_profile
        if prof and prof.tournament_max_rounds_overrides:
            kw = dict(target_score=prof.default_target_score, max_rounds=prof.default_max_rounds)

            def launch(root_seed, kk, s0, n, table, tallies, totals):
                res = eng.play_tournament(root_seed, kk, s0, n, table, tallies=tallies, totals=totals,
                                          overrides=prof.tournament_overrides_for(root_seed, kk, s0, n),
                                          want_first_seen=True, **kw)
                return res.tallies, res.totals, res.first_seen
        else:
            kw = dict(target_score=prof.default_target_score, max_rounds=prof.default_max_rounds) if prof else {}
            launch = _engine_launch(eng, kw)
        table_dev = eng.to_device(state.table)
        tallies, totals, seen = run_cell(root, k, cfg.num_shuffles, table_dev, batch_size=batch, launch=launch,
                                         rank=rank, world=world, want_first_seen=True)
        if torch.device(eng.device).type == "cuda":
            torch.cuda.synchronize(eng.device)
        if rank != 0:
            return
        tot = totals.cpu().numpy()
        if tot[7]:
            raise RuntimeError("a game hit ROLL_LIMIT or overflowed an int16 row counter")
        wins, sums, sqs = tallies_to_outcome(tallies.cpu().numpy()[0], state.ids, tuple(tot[:3]),
                                             seen.cpu().numpy())
    else:
        # ---- batched route: groups of whole deterministic batches, in order
        wins, sums, sqs = OutcomeCounter(), {m: defaultdict(float) for m in METRIC_LABELS}, \
            {m: defaultdict(float) for m in METRIC_LABELS}
        mine = shard_batches(cfg.num_shuffles, batch, rank, world)
        per_group = 1 if chunk_dir else ROWS_LAUNCH_BATCHES
        newly: set[int] = set()
        last_ckpt = time.monotonic()
        for g0 in range(0, len(mine), per_group):
            group = mine[g0:g0 + per_group]
            shuffles = [s for _, s0, cnt in group for s in range(s0, s0 + cnt)]
            count = [s for s in shuffles if s not in completed]           # tallies still owed
            need_rows = [s for s in shuffles if collect_rows and s not in rows_done]
            both = [s for s in count if s in set(need_rows)]
            only_tallies = [s for s in count if s not in set(need_rows)]
            only_rows = [s for s in need_rows if s in completed]
            part_w, part_s, part_q = OutcomeCounter(), {m: defaultdict(float) for m in METRIC_LABELS}, \
                {m: defaultdict(float) for m in METRIC_LABELS}
            for todo, rows_on, counted in ((both, True, True), (only_tallies, False, True),
                                           (only_rows, True, False)):
                if not todo:
                    continue
                w, sm, sq = _run_chunk_metrics(make_shuffle_tasks(root, k, todo, batch), collect_rows=rows_on,
                                               row_dir=row_dir, manifest_path=row_manifest, io_threads=io_threads)
                if counted:
                    part_w.absorb(w)
                    _add_sums(part_s, sm)
                    _add_sums(part_q, sq)
            if chunk_dir and count and (group[0][0] + 1) not in chunks_done:
                _write_metric_chunk(chunk_dir, state, root, make_shuffle_tasks(root, k, count, batch),
                                    part_w, part_s, part_q)
            wins.absorb(part_w)
            _add_sums(sums, part_s)
            _add_sums(sqs, part_q)
            newly |= set(count)
            if world == 1 and time.monotonic() - last_ckpt >= cfg.ckpt_every_sec:
                done_now = completed | newly
                w_now = OutcomeCounter()
                w_now.absorb(wins_prev)
                w_now.absorb(wins)
                s_now = {m: defaultdict(float, (sums_prev or {}).get(m, {})) for m in METRIC_LABELS}
                q_now = {m: defaultdict(float, (sqs_prev or {}).get(m, {})) for m in METRIC_LABELS}
                _add_sums(s_now, sums)
                _add_sums(q_now, sqs)
                _save_checkpoint(ckpt_path, w_now, s_now if want_sums else None, q_now if want_sums else None,
                                 {**meta, "completed_shuffle_indices": sorted(done_now),
                                  "completed_process_block_indices": sorted({s // batch + 1 for s in done_now})})
                last_ckpt = time.monotonic()
        if world > 1:   # join the ranks' partial counters in rank order (= batch order)
            parts: List[Any] = [None] * world
            dist.all_gather_object(parts, (wins, sums, sqs, sorted(newly)))
            wins, sums, sqs, newly = OutcomeCounter(), {m: defaultdict(float) for m in METRIC_LABELS}, \
                {m: defaultdict(float) for m in METRIC_LABELS}, set()
            for w, sm, sq, done in parts:
                wins.absorb(w)
                _add_sums(sums, sm)
                _add_sums(sqs, sq)
                newly |= set(done)
        if rank != 0:
            return
        if chunk_dir and sums_prev is None and (chunk_dir / "metrics_manifest.jsonl").exists():
            # no checkpointed sums to add to: the reference takes the run's aggregates from the
            # chunk shards themselves (run_tournament.py:1770-1784)
            wins, sums, sqs = _load_metric_chunk_aggregates(chunk_dir / "metrics_manifest.jsonl", k)
            wins_prev = OutcomeCounter()
        total_w = OutcomeCounter()
        total_w.absorb(wins_prev)
        total_w.absorb(wins)
        wins = total_w
        for prev, cur in ((sums_prev, sums), (sqs_prev, sqs)):
            if prev is not None:
                for label in METRIC_LABELS:          # earlier aggregates first: the reference's key order
                    merged = defaultdict(float, prev[label])
                    for key, v in cur[label].items():
                        merged[key] += v
                    cur[label] = merged
        completed |= newly
    if plain:
        completed = set(range(cfg.num_shuffles))
    meta["completed_shuffle_indices"] = sorted(completed)
    meta["completed_process_block_indices"] = sorted({s // batch + 1 for s in completed})
    del n_batches
    _save_checkpoint(ckpt_path, wins, sums if want_sums else None, sqs if want_sums else None, meta)
    if write_final_metrics_artifact and want_sums:
        import pyarrow as pa
        import pyarrow.parquet as pq

        # label-major rows for the strategies that won at least once (run_tournament.py:1797-1826)
        labels, strategies_col, sum_col, sq_col = [], [], [], []
        for label in METRIC_LABELS:
            for strat, value in sums[label].items():
                labels.append(label)
                strategies_col.append(int(strat))
                sum_col.append(float(value))
                sq_col.append(float(sqs[label].get(strat, 0.0)))
        schema = pa.schema([pa.field("metric", pa.string()), pa.field("strategy", pa.int32()),
                            pa.field("sum", pa.float64()), pa.field("square_sum", pa.float64())])
        tbl = pa.Table.from_arrays([pa.array(labels, type=pa.string()),
                                    pa.array(np.asarray(strategies_col, dtype=np.int32)),
                                    pa.array(np.asarray(sum_col, dtype=np.float64)),
                                    pa.array(np.asarray(sq_col, dtype=np.float64))], schema=schema)
        _atomic_write(ckpt_path.with_name(f"{k}p_metrics.parquet"), lambda p: pq.write_table(tbl, p))
    LOGGER.info("Tournament run complete after %d attempted games", wins.games_attempted)


# --------------------------------------------------------------------------- seat tallies
def seat_counts_from_rows(rows: np.ndarray, batch_ids: np.ndarray) -> Dict[Tuple[int, int, int], List[int]]:
    """``{(batch, strategy, seat 1-based): [wins, exposures, completed, safety_limit]}`` from compact
    rows — the counts of the reference's ``_iter_seat_count_tables`` (analysis/seat_analysis.py:166-229).
    Host restatement used to check the device-side seat tallies."""
    out: Dict[Tuple[int, int, int], List[int]] = {}
    k = rows["seats"].shape[1]
    safety = (rows["flags"] & 1) != 0
    for s in range(k):
        strat = rows["seats"]["strategy"][:, s]
        won = (~safety) & (rows["winner_seat"] == s)
        for b, sid, w, sl in zip(batch_ids.tolist(), strat.tolist(), won.tolist(), safety.tolist()):
            cell = out.setdefault((b, sid, s + 1), [0, 0, 0, 0])
            cell[0] += int(w)
            cell[1] += 1
            cell[2] += int(not sl)
            cell[3] += int(sl)
    return out


def seat_counts_table(seat_tallies: np.ndarray, ids: Sequence[int], *, root_seed: int, k: int,
                      first_batch_id: int = 0):
    """Device seat tallies ``[slots, ids, k, 4]`` -> Arrow table with the reference's seat-count
    schema (analysis/seat_analysis.py:42-54), rows ordered by (batch, strategy, seat), empty cells
    dropped."""
    import pyarrow as pa

    t = np.asarray(seat_tallies)
    n_slots, n_ids = t.shape[0], t.shape[1]
    order = np.argsort(np.asarray(ids), kind="stable")
    t = t[:, order]
    sid = np.asarray(ids)[order]
    slot_i, id_i, seat_i = np.nonzero(t[..., 1])
    schema = pa.schema([
        pa.field("root_seed", pa.int64(), nullable=False), pa.field("k", pa.int16(), nullable=False),
        pa.field("deterministic_batch_id", pa.int32(), nullable=False),
        pa.field("strategy", pa.int32(), nullable=False), pa.field("seat", pa.int16(), nullable=False),
        pa.field("raw_wins", pa.int64(), nullable=False),
        pa.field("raw_exposures", pa.int64(), nullable=False),
        pa.field("raw_completed_exposures", pa.int64(), nullable=False),
        pa.field("raw_safety_limit_exposures", pa.int64(), nullable=False)])
    cells = t[slot_i, id_i, seat_i]
    n = len(slot_i)
    del n_slots, n_ids
    return pa.Table.from_arrays([
        pa.array(np.full(n, root_seed, dtype=np.int64)), pa.array(np.full(n, k, dtype=np.int16)),
        pa.array((slot_i + first_batch_id).astype(np.int32)), pa.array(sid[id_i].astype(np.int32)),
        pa.array((seat_i + 1).astype(np.int16)), pa.array(cells[:, 0]), pa.array(cells[:, 1]),
        pa.array(cells[:, 2]), pa.array(cells[:, 3])], schema=schema)


# --------------------------------------------------------------------------- all-player statistics
ALL_PLAYER_BEHAVIOURS = ("rank", "loss_margin", "rolls", "farkles", "highest_turn", "hot_dice",
                         "smart_five_uses", "n_smart_five_dice", "smart_one_uses", "n_smart_one_dice")
_ALLP_COUNTS = ("raw_player_game_exposures", "raw_completed_player_game_exposures",
                "raw_safety_limit_player_game_exposures", "raw_wins", "raw_losses",
                "raw_turn_round_mismatch_count", "raw_max_round_abort_exposures")
_ALLP_SUMS = ("raw_final_score_sum", "raw_final_score_square_sum", "raw_n_turns_sum", "raw_n_turns_square_sum",
              "raw_turn_return_game_weighted_exact_sum", "raw_turn_return_game_weighted_exact_square_sum",
              "raw_turn_return_round_proxy_sum", "raw_turn_return_round_proxy_square_sum",
              "raw_turn_minus_rounds_sum", "raw_turn_minus_rounds_square_sum")
_ALLP_DERIVED = ("turn_return_turn_weighted", "turn_return_game_weighted_exact", "turn_return_round_proxy",
                 "round_proxy_gap", "round_proxy_relative_gap", "turn_round_mismatch_prevalence",
                 "win_rate_per_attempt", "win_rate_given_completion", "safety_limit_exposure_rate")


def all_player_schema():
    """The reference's ``all_player_batch_schema()`` (analysis/all_player_metrics.py:101-121)."""
    import pyarrow as pa

    behaviour = []
    for suffix in ALL_PLAYER_BEHAVIOURS:
        behaviour += [pa.field(f"raw_{suffix}_observations", pa.int64(), nullable=False),
                      pa.field(f"raw_{suffix}_sum", pa.float64(), nullable=False),
                      pa.field(f"raw_{suffix}_square_sum", pa.float64(), nullable=False)]
    return pa.schema([
        pa.field("root_seed", pa.int64(), nullable=False), pa.field("k", pa.int16(), nullable=False),
        pa.field("deterministic_batch_id", pa.int32(), nullable=False),
        pa.field("strategy", pa.int32(), nullable=False),
        *(pa.field(name, pa.int64(), nullable=False) for name in _ALLP_COUNTS),
        *(pa.field(name, pa.float64(), nullable=False) for name in _ALLP_SUMS),
        *behaviour, *(pa.field(name, pa.float64()) for name in _ALLP_DERIVED)])


def all_player_table(all_player: np.ndarray, ids: Sequence[int], *, root_seed: int, k: int,
                     first_batch_id: int = 0):
    """Device all-player statistics ``int64 [slots, ids, ALLP_WIDTH]`` (``Engine.play_tournament(
    want_all_player=True, shuffles_per_slot=batch)``) -> the Arrow table the reference's metrics
    stage builds by re-reading every curated row (analysis/all_player_metrics.py:425-520): one row
    per (batch, strategy) in (batch, strategy) order, the raw sufficient statistics and the derived
    ratios computed exactly as ``_finish_row`` (:384-423) computes them."""
    import pyarrow as pa

    t = np.ascontiguousarray(all_player, dtype=np.int64)
    order = np.argsort(np.asarray(ids), kind="stable")
    sid = np.asarray(ids)[order]
    rows = []
    for slot in range(t.shape[0]):
        for pos, strategy in zip(order.tolist(), sid.tolist()):
            v = t[slot, pos]
            exposures, completed, safety, wins = (int(x) for x in v[:4])
            if exposures == 0:
                continue
            f = v[41:45].view(np.float64)
            exact_sum, exact_sq, proxy_sum, proxy_sq = (float(x) for x in f)
            turns = float(v[7])
            game_exact = exact_sum / exposures
            round_proxy = proxy_sum / exposures
            gap = round_proxy - game_exact
            row: Dict[str, Any] = {
                "root_seed": root_seed, "k": k, "deterministic_batch_id": first_batch_id + slot,
                "strategy": int(strategy),
                "raw_player_game_exposures": exposures, "raw_completed_player_game_exposures": completed,
                "raw_safety_limit_player_game_exposures": safety, "raw_wins": wins,
                "raw_losses": exposures - wins, "raw_turn_round_mismatch_count": int(v[4]),
                "raw_max_round_abort_exposures": safety,
                "raw_final_score_sum": float(v[5]), "raw_final_score_square_sum": float(v[6]),
                "raw_n_turns_sum": turns, "raw_n_turns_square_sum": float(v[8]),
                "raw_turn_return_game_weighted_exact_sum": exact_sum,
                "raw_turn_return_game_weighted_exact_square_sum": exact_sq,
                "raw_turn_return_round_proxy_sum": proxy_sum,
                "raw_turn_return_round_proxy_square_sum": proxy_sq,
                "raw_turn_minus_rounds_sum": float(v[9]), "raw_turn_minus_rounds_square_sum": float(v[10]),
                "turn_return_turn_weighted": float(v[5]) / turns if turns else None,
                "turn_return_game_weighted_exact": game_exact, "turn_return_round_proxy": round_proxy,
                "round_proxy_gap": gap, "round_proxy_relative_gap": gap / game_exact if game_exact else None,
                "turn_round_mismatch_prevalence": float(v[4]) / exposures,
                "win_rate_per_attempt": wins / exposures,
                "win_rate_given_completion": wins / completed if completed else None,
                "safety_limit_exposure_rate": safety / exposures}
            for b, suffix in enumerate(ALL_PLAYER_BEHAVIOURS):
                row[f"raw_{suffix}_observations"] = int(v[11 + 3 * b])
                row[f"raw_{suffix}_sum"] = float(v[12 + 3 * b])
                row[f"raw_{suffix}_square_sum"] = float(v[13 + 3 * b])
            rows.append(row)
    return pa.Table.from_pylist(rows, schema=all_player_schema())
