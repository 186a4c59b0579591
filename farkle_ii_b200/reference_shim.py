"""Plug the CUDA engine into an unmodified Farkle_II checkout.

``install()`` replaces, inside the reference's own ``farkle.simulation.run_tournament`` module, the
worker callables its driver dispatches to — the same seams the reference's tests patch
(``tests/unit/simulation/test_run_tournament.py:45-63``):

    _init_worker        run_tournament.py:265
    _play_one_shuffle   run_tournament.py:301
    _play_shuffle       run_tournament.py:396
    _run_chunk          run_tournament.py:403
    _run_chunk_metrics  run_tournament.py:473

Everything above them (``run_tournament``'s chunk scheduling, checkpoints, metric chunks, resume,
``farkle run``) keeps running the reference's code; with ``n_jobs=1`` the "worker" is the GPU of
this process.  Results travel back in the reference's own types (its ``OutcomeCounter``, plain
``dict[label][strategy] -> float``, row dicts).
"""

from __future__ import annotations

from collections import defaultdict
from pathlib import Path
from typing import Any, Sequence

from . import run_tournament as gpu_rt
from .limits import GameProfile, H2HMaxRoundsOverride, TournamentMaxRoundsOverride
from .strategies import FavorDiceOrScore, ThresholdStrategy

_SEAMS = ("_init_worker", "_play_one_shuffle", "_play_shuffle", "_run_chunk", "_run_chunk_metrics")


def _to_gpu_strategy(s: Any) -> ThresholdStrategy:
    favor = s.favor_dice_or_score
    favor = FavorDiceOrScore(getattr(favor, "value", favor))
    return ThresholdStrategy(
        score_threshold=int(s.score_threshold), dice_threshold=int(s.dice_threshold),
        smart_five=bool(s.smart_five), smart_one=bool(s.smart_one),
        consider_score=bool(s.consider_score), consider_dice=bool(s.consider_dice),
        require_both=bool(s.require_both), auto_hot_dice=bool(s.auto_hot_dice),
        run_up_score=bool(s.run_up_score), favor_dice_or_score=favor,
        strategy_id=getattr(s, "strategy_id", None))


def _to_gpu_profile(p: Any) -> GameProfile | None:
    if p is None:
        return None
    return GameProfile(
        default_target_score=p.default_target_score, default_max_rounds=p.default_max_rounds,
        tournament_max_rounds_overrides=tuple(
            TournamentMaxRoundsOverride(o.root_seed, o.k, o.shuffle_index, o.game_index, o.max_rounds)
            for o in p.tournament_max_rounds_overrides),
        h2h_max_rounds_overrides=tuple(
            H2HMaxRoundsOverride(o.root_seed, o.pair_id, o.order, o.attempt_index, o.max_rounds)
            for o in p.h2h_max_rounds_overrides))


def install(rt: Any = None, *, device: int | None = None) -> dict[str, Any]:
    """Patch the reference module ``rt`` (default: ``farkle.simulation.run_tournament``).

    Returns the original callables; pass them to :func:`uninstall` to restore.
    """
    if rt is None:
        import farkle.simulation.run_tournament as rt  # type: ignore[no-redef]
    originals = {name: getattr(rt, name) for name in _SEAMS}

    def task_of(task: Any) -> gpu_rt.ShuffleTask:
        work = rt._coerce_shuffle_task(task)
        return gpu_rt.ShuffleTask(work.root_seed, work.k, work.shuffle_index, work.shuffle_seed,
                                  work.deterministic_batch_id)

    def counter_of(w: gpu_rt.OutcomeCounter) -> Any:
        c = rt.OutcomeCounter(dict(w))
        c.attempted_exposures.update(w.attempted_exposures)
        c.completed_exposures.update(w.completed_exposures)
        c.safety_limit_exposures.update(w.safety_limit_exposures)
        c.games_attempted, c.games_completed = w.games_attempted, w.games_completed
        c.games_safety_limit = w.games_safety_limit
        return c

    def sums_of(s: dict) -> dict:
        return {label: defaultdict(float, v) for label, v in s.items()}

    def _init_worker(strategies: Sequence[Any], config: Any, game_profile: Any = None,
                     progress_endpoint: Any = None) -> None:
        originals["_init_worker"](strategies, config, game_profile, progress_endpoint)
        state = rt._STATE  # the reference resolved the strategy ids; mirror exactly that list
        cfg = gpu_rt.TournamentConfig(
            n_players=config.n_players, num_shuffles=config.num_shuffles,
            n_strategies=len(state.strats),
            deterministic_batch_size=getattr(config, "deterministic_batch_size", 30))
        gpu_rt._init_worker([_to_gpu_strategy(s) for s in state.strats], cfg,
                            _to_gpu_profile(game_profile), device=device)

    def _play_one_shuffle(task: Any, *, collect_rows: bool = False):
        w, s, q, rows = gpu_rt._play_one_shuffle(task_of(task), collect_rows=collect_rows)
        return counter_of(w), sums_of(s), sums_of(q), rows

    def _play_shuffle(task: Any):
        return _play_one_shuffle(task, collect_rows=False)[0]

    def _run_chunk(shuffle_tasks: Sequence[Any]):
        return counter_of(gpu_rt._run_chunk([task_of(t) for t in shuffle_tasks]))

    def _run_chunk_metrics(shuffle_tasks: Sequence[Any], *, collect_rows: bool = False,
                           row_dir: Path | None = None, manifest_path: Path | None = None,
                           row_sidecar: Any = None):
        def reference_writer(out: Path, manifest_file: Path, table: Any, extra: Any) -> None:
            # the reference's own shard publisher: Parquet writer, manifest line and — when the
            # runner passes one — the hash-bound sidecar (run_tournament.py:530-558)
            rt.run_streaming_shard(out_path=str(out), manifest_path=str(manifest_file),
                                   schema=table.schema, batch_iter=(table,),
                                   manifest_extra=dict(extra), sidecar=row_sidecar)

        w, s, q = gpu_rt._run_chunk_metrics([task_of(t) for t in shuffle_tasks],
                                            collect_rows=collect_rows, row_dir=row_dir,
                                            manifest_path=manifest_path, row_sidecar=row_sidecar,
                                            shard_writer=reference_writer)
        return counter_of(w), sums_of(s), sums_of(q)

    for name, fn in (("_init_worker", _init_worker), ("_play_one_shuffle", _play_one_shuffle),
                     ("_play_shuffle", _play_shuffle), ("_run_chunk", _run_chunk),
                     ("_run_chunk_metrics", _run_chunk_metrics)):
        setattr(rt, name, fn)
    return originals


def uninstall(originals: dict[str, Any], rt: Any = None) -> None:
    if rt is None:
        import farkle.simulation.run_tournament as rt  # type: ignore[no-redef]
    for name, fn in originals.items():
        setattr(rt, name, fn)
