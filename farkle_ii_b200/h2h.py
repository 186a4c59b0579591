"""Head-to-head block execution over the CUDA engine.

Mirror of the reference's H2H seam (src/farkle/analysis/h2h_schedule.py):

* ``_block_progress``                    :1088-1146  (result dict, same keys)
* ``_simulate_block_from_manifest``      :1149-1243  (attempt loop with early stop)
* ``BlockRunner``                        :1521       ``(block, strategy_manifest, chunk_games) -> dict``

A block's attempts are independent coordinates ``(root, pair_id, order, attempt_index)``, so
instead of playing them one by one the runner launches exactly as many attempts as are
still *needed* (``n_completed_required - games_completed``, capped by ``chunk_games`` and
``max_attempts``), lets ``fb_h2h_resolve`` apply the reference's prefix early-stop rule, and
repeats for the few attempts lost to the safety limit.  ``simulate_blocks`` does this for
many blocks per launch; the single-block ``BlockRunner`` is a thin wrapper.
"""

from __future__ import annotations

import hashlib
import json
from pathlib import Path
from typing import Any, Mapping, Sequence

import numpy as np

from .limits import GameProfile
from .layout import ROW_I16_OVERFLOW, ROW_ROLL_LIMIT
from .random import RandomPurpose
from .simulation import PlayerRngCoordinates, RollLimitError, play_games_batch
from .strategies import (
    STRATEGY_TUPLE_FIELDS,
    FavorDiceOrScore,
    ThresholdStrategy,
    pack_strategies,
)

PARTIAL_RESUMABLE = "partial_resumable"

_PROGRESS_KEYS = {
    "games_attempted", "games_completed", "games_safety_limit", "wins_seat1", "wins_seat2",
    "wins_a", "wins_b", "replacement_attempt_count", "completion_status", "completion_game_rate",
    "safety_limit_game_rate", "authenticated_attempt_index_start",
    "authenticated_attempt_index_stop_exclusive", "attempt_coordinate_range_hash",
}


def _attempt_coordinate_range_hash(block: Mapping[str, Any], stop_exclusive: int) -> str:
    """h2h_schedule.py:1072-1085."""
    payload = {
        "rng_scheme_version": int(block["rng_scheme_version"]),
        "purpose": int(block["rng_purpose_namespace"]),
        "root_seed": int(block["root_seed"]),
        "pair_id": int(block["pair_id"]),
        "order": int(block["order"]),
        "attempt_index_start": 0,
        "attempt_index_stop_exclusive": int(stop_exclusive),
    }
    return hashlib.sha256(
        json.dumps(payload, sort_keys=True, separators=(",", ":")).encode("utf-8")).hexdigest()


def _block_progress(block: Mapping[str, Any], *, games_attempted: int, games_completed: int,
                    games_safety_limit: int, wins_seat1: int, wins_seat2: int) -> dict[str, Any]:
    """Result dict of one block advance (h2h_schedule.py:1088-1146)."""
    target = int(block["n_completed_required"])
    max_attempts = int(block["max_attempts"])
    if games_completed >= target:
        status = "complete"
    elif games_attempted >= max_attempts:
        status = "unresolved_nonviable"
    else:
        status = PARTIAL_RESUMABLE
    out = {key: value for key, value in block.items()
           if not str(key).startswith("_") and key not in _PROGRESS_KEYS}
    out.update({
        "wins_a": wins_seat1 if int(block["order"]) == 0 else wins_seat2,
        "wins_b": wins_seat2 if int(block["order"]) == 0 else wins_seat1,
        "games_attempted": games_attempted,
        "games_completed": games_completed,
        "games_safety_limit": games_safety_limit,
        "wins_seat1": wins_seat1,
        "wins_seat2": wins_seat2,
        "replacement_attempt_count": max(0, games_attempted - target),
        "completion_status": status,
        "completion_game_rate": games_completed / games_attempted if games_attempted else None,
        "safety_limit_game_rate": games_safety_limit / games_attempted if games_attempted else None,
        "authenticated_attempt_index_start": 0,
        "authenticated_attempt_index_stop_exclusive": games_attempted,
    })
    if "rng_scheme_version" in block and "rng_purpose_namespace" in block:
        out["attempt_coordinate_range_hash"] = _attempt_coordinate_range_hash(block, games_attempted)
    return out


def parse_strategy_identifier(value: Any, manifest: Any) -> ThresholdStrategy:
    """Canonical numeric id -> strategy (simulation/strategies.py:762-800).

    ``manifest`` is the strategy manifest DataFrame (columns ``strategy_id`` + the ten
    strategy fields) or a mapping ``id -> ThresholdStrategy``.
    """
    if (isinstance(value, (int, np.integer)) and not isinstance(value, bool)) or (
            isinstance(value, str) and value.isdigit()):
        sid = int(value)
    else:
        raise ValueError(f"Cannot parse nonnumeric strategy identifier: {value!r}")
    if isinstance(manifest, Mapping):
        if sid not in manifest:
            raise KeyError(f"strategy_id {sid} missing from manifest/encoder")
        s = manifest[sid]
        return s if s.strategy_id == sid else ThresholdStrategy(
            **{f: getattr(s, f) for f in STRATEGY_TUPLE_FIELDS}, strategy_id=sid)
    match = manifest.loc[manifest["strategy_id"] == sid]
    if match.empty:
        raise KeyError(f"strategy_id {sid} missing from manifest/encoder")
    attrs = {str(k): v for k, v in match.iloc[0].to_dict().items() if k in STRATEGY_TUPLE_FIELDS}
    favor = attrs.get("favor_dice_or_score")
    if favor is not None and not isinstance(favor, FavorDiceOrScore):
        attrs["favor_dice_or_score"] = (FavorDiceOrScore.SCORE if favor == FavorDiceOrScore.SCORE.value
                                        else FavorDiceOrScore.DICE)
    for name in ("score_threshold", "dice_threshold"):
        attrs[name] = int(attrs[name])
    for name in ("smart_five", "smart_one", "consider_score", "consider_dice", "require_both",
                 "auto_hot_dice", "run_up_score"):
        attrs[name] = bool(attrs[name])
    return ThresholdStrategy(**attrs, strategy_id=sid)


_PACKED_MANIFESTS: dict[int, tuple[Any, np.ndarray, np.ndarray]] = {}


def _packed_manifest(manifest: Any) -> tuple[np.ndarray, np.ndarray]:
    """``(sorted strategy ids, packed fb_strategy_t rows)`` of a manifest, built once per manifest
    object (a schedule looks the same two columns up for tens of thousands of blocks)."""
    from .layout import (SF_AUTO_HOT_DICE, SF_CONSIDER_DICE, SF_CONSIDER_SCORE, SF_FAVOR_SCORE,
                         SF_REQUIRE_BOTH, SF_RUN_UP_SCORE, SF_SMART_FIVE, SF_SMART_ONE, STRATEGY_DTYPE)

    hit = _PACKED_MANIFESTS.get(id(manifest))
    if hit is not None and hit[0] is manifest:
        return hit[1], hit[2]
    if isinstance(manifest, Mapping):
        ids = np.array(sorted(manifest), dtype=np.int64)
        packed = pack_strategies([manifest[int(i)] for i in ids])
    else:
        frame = manifest.sort_values("strategy_id", kind="mergesort")
        ids = frame["strategy_id"].to_numpy(dtype=np.int64)
        flags = np.zeros(len(frame), dtype=np.uint16)
        for col, bit in (("smart_five", SF_SMART_FIVE), ("smart_one", SF_SMART_ONE),
                         ("consider_score", SF_CONSIDER_SCORE), ("consider_dice", SF_CONSIDER_DICE),
                         ("require_both", SF_REQUIRE_BOTH), ("auto_hot_dice", SF_AUTO_HOT_DICE),
                         ("run_up_score", SF_RUN_UP_SCORE)):
            flags |= np.where(frame[col].to_numpy(dtype=bool), bit, 0).astype(np.uint16)
        favor = frame["favor_dice_or_score"].map(
            lambda v: (v if isinstance(v, str) else getattr(v, "value", v)) == FavorDiceOrScore.SCORE.value)
        flags |= np.where(favor.to_numpy(dtype=bool), SF_FAVOR_SCORE, 0).astype(np.uint16)
        smart_one_without_five = (flags & SF_SMART_ONE != 0) & (flags & SF_SMART_FIVE == 0)
        both_without_two = (flags & SF_REQUIRE_BOTH != 0) & (
            (flags & SF_CONSIDER_SCORE == 0) | (flags & SF_CONSIDER_DICE == 0))
        if smart_one_without_five.any() or both_without_two.any():
            raise ValueError("strategy manifest holds a combination ThresholdStrategy rejects")
        packed = np.zeros(len(frame), dtype=STRATEGY_DTYPE)
        packed["score_threshold"] = frame["score_threshold"].to_numpy(dtype=np.int64)
        packed["dice_threshold"] = frame["dice_threshold"].to_numpy(dtype=np.int64)
        packed["flags"] = flags
    if len(_PACKED_MANIFESTS) > 4:
        _PACKED_MANIFESTS.clear()
    _PACKED_MANIFESTS[id(manifest)] = (manifest, ids, packed)
    return ids, packed


def _lookup_packed(manifest: Any, wanted: Sequence[Any]) -> np.ndarray:
    """Packed strategies of canonical numeric identifiers (strategies.py:762-800), vectorised."""
    numeric = []
    for value in wanted:
        if (isinstance(value, (int, np.integer)) and not isinstance(value, bool)) or (
                isinstance(value, str) and value.isdigit()):
            numeric.append(int(value))
        else:
            raise ValueError(f"Cannot parse nonnumeric strategy identifier: {value!r}")
    ids, packed = _packed_manifest(manifest)
    want = np.asarray(numeric, dtype=np.int64)
    pos = np.searchsorted(ids, want)
    pos_c = np.minimum(pos, max(len(ids) - 1, 0))
    missing = (len(ids) == 0) | (ids[pos_c] != want) if len(ids) else np.ones(len(want), dtype=bool)
    if np.any(missing):
        raise KeyError(f"strategy_id {int(want[np.argmax(missing)])} missing from manifest/encoder")
    return packed[pos_c]


def _check_outcomes(outcome: np.ndarray) -> None:
    if (outcome & 0x80).any():
        raise RollLimitError("an H2H attempt hit ROLL_LIMIT or overflowed an int16 row counter")


def simulate_blocks(blocks: Sequence[Mapping[str, Any]], manifest: Any, chunk_games: int,
                    oracle_game_profile: GameProfile | None = None, *,
                    device: int | None = None) -> list[dict[str, Any]]:
    """Advance every block by at most ``chunk_games`` attempts — many blocks per launch.

    Equal, block for block, to ``[_simulate_block_from_manifest(b, manifest, chunk_games,
    oracle_game_profile) for b in blocks]`` of the reference.  All blocks must share one
    ``root_seed`` per launch group (they are grouped here).
    """
    from .device import get_engine

    n = len(blocks)
    if n == 0:
        return []
    if chunk_games < 0:
        raise ValueError("chunk_games must be non-negative")
    prof = oracle_game_profile
    target_score = prof.default_target_score if prof else 10_000
    max_rounds = prof.default_max_rounds if prof else 200
    s1 = _lookup_packed(manifest, [b["seat1_strategy"] for b in blocks])
    s2 = _lookup_packed(manifest, [b["seat2_strategy"] for b in blocks])
    root = np.array([int(b["root_seed"]) for b in blocks], dtype=np.uint64)
    pair = np.array([int(b["pair_id"]) for b in blocks], dtype=np.uint64)
    order = np.array([int(b["order"]) for b in blocks], dtype=np.uint8)
    target = np.array([int(b["n_completed_required"]) for b in blocks], dtype=np.int64)
    max_attempts = np.array([int(b["max_attempts"]) for b in blocks], dtype=np.int64)
    prog = np.array([[int(b.get(key, 0)) for key in ("games_attempted", "games_completed",
                                                     "games_safety_limit", "wins_seat1", "wins_seat2")]
                     for b in blocks], dtype=np.int64)
    stop = np.minimum(max_attempts, prog[:, 0] + chunk_games)
    overridden = {(o.root_seed, o.pair_id, o.order): True
                  for o in (prof.h2h_max_rounds_overrides if prof else ())}
    eng = get_engine(device)
    while True:
        need = np.minimum(np.maximum(target - prog[:, 1], 0), np.maximum(stop - prog[:, 0], 0))
        active = np.flatnonzero(need > 0)
        if len(active) == 0:
            break
        for r in np.unique(root[active]):
            idx = active[root[active] == r]
            plain = np.array([i for i in idx
                              if (int(r), int(pair[i]), int(order[i])) not in overridden], dtype=np.int64)
            if len(plain):
                outcome, d_na, _rows, _tot = eng.play_h2h(
                    int(r), pair[plain], order[plain], s1[plain], s2[plain],
                    prog[plain, 0].astype(np.uint32), need[plain].astype(np.uint32),
                    target_score=target_score, max_rounds=max_rounds)
                _check_outcomes(outcome.cpu().numpy())
                prog[plain] = eng.h2h_resolve(d_na, outcome, target[plain].astype(np.int32),
                                              prog[plain].astype(np.int32))
            for i in idx:  # blocks with per-attempt max_rounds overrides: explicit coordinates
                if (int(r), int(pair[i]), int(order[i])) not in overridden:
                    continue
                _advance_with_overrides(blocks[i], manifest, prof, prog, i, int(need[i]), device)
    keys = ("games_attempted", "games_completed", "games_safety_limit", "wins_seat1", "wins_seat2")
    return [_block_progress(b, **{key: int(prog[i, j]) for j, key in enumerate(keys)})
            for i, b in enumerate(blocks)]


def _advance_with_overrides(block, manifest, prof: GameProfile, prog: np.ndarray, i: int, need: int,
                            device: int | None) -> None:
    """Attempts of a block whose coordinates carry max_rounds overrides (game_profile.py:182-200)."""
    st = [parse_strategy_identifier(block["seat1_strategy"], manifest),
          parse_strategy_identifier(block["seat2_strategy"], manifest)]
    if st[0].strategy_id == st[1].strategy_id:
        raise ValueError("Simulation row must seat distinct strategies")
    a0 = int(prog[i, 0])
    root, pair_id, order = int(block["root_seed"]), int(block["pair_id"]), int(block["order"])
    coords = [PlayerRngCoordinates(purpose=RandomPurpose.H2H_PLAYER, root_seed=root, k=2,
                                   pair_id=pair_id, order=order, attempt_index=a)
              for a in range(a0, a0 + need)]
    limits = [prof.h2h_limits(root_seed=root, pair_id=pair_id, order=order, attempt_index=a)
              for a in range(a0, a0 + need)]
    rows = play_games_batch([st] * need, coords, target_score=prof.default_target_score,
                            max_rounds=[lim.max_rounds for lim in limits], device=device)
    if (rows["flags"] & (ROW_ROLL_LIMIT | ROW_I16_OVERFLOW)).any():
        raise RollLimitError("an H2H attempt hit ROLL_LIMIT or overflowed an int16 row counter")
    target = int(block["n_completed_required"])
    for r in rows:
        if prog[i, 1] >= target:
            break
        prog[i, 0] += 1
        if r["flags"] & 1:
            prog[i, 2] += 1
        else:
            prog[i, 1] += 1
            prog[i, 3 + int(r["winner_seat"])] += 1


def _simulate_block_from_manifest(block: dict[str, Any], manifest: Any, chunk_games: int,
                                  oracle_game_profile: GameProfile | None = None, *,
                                  device: int | None = None) -> dict[str, Any]:
    """Advance one root/order block (h2h_schedule.py:1149-1243)."""
    return simulate_blocks([block], manifest, chunk_games, oracle_game_profile, device=device)[0]


_MANIFEST_CACHE: dict[tuple[str, int, int], Any] = {}


def _load_manifest(path: Path):
    import pandas as pd

    st = Path(path).stat()
    key = (str(Path(path).resolve()), st.st_mtime_ns, st.st_size)
    if key not in _MANIFEST_CACHE:
        _MANIFEST_CACHE.clear()
        _MANIFEST_CACHE[key] = pd.read_parquet(path)
    return _MANIFEST_CACHE[key]


def gpu_block_runner(block: dict[str, Any], strategy_manifest: Path, chunk_games: int) -> dict[str, Any]:
    """``BlockRunner`` (h2h_schedule.py:1521) backed by the CUDA engine."""
    return _simulate_block_from_manifest(block, _load_manifest(strategy_manifest), chunk_games)


class BatchedBlockRunner:
    """A ``BlockRunner`` for the UNMODIFIED ``execute_h2h_schedule(cfg, block_runner=...)`` that
    still plays many blocks per launch.

    The stage calls its runner one block at a time (h2h_schedule.py:2040-2061), which would cost a
    launch-resolve round trip per block.  Built with the blocks the stage is about to ask for
    (its pending schedule rows, e.g. ``pd.read_parquet(cfg.h2h_block_manifest_path())`` as dicts)
    and the stage's ``chunk_games``, this runner answers the first call by advancing ALL of them
    together -- every chunk of every block until each is terminal -- and serves the later calls from
    that table.  A call it did not foresee (unknown block, different progress or chunk bound) is
    simply simulated on its own, so the answers are those of ``gpu_block_runner`` in every case.
    """

    def __init__(self, blocks: Sequence[Mapping[str, Any]], *, chunk_games: int = 5_000,
                 oracle_game_profile: GameProfile | None = None, device: int | None = None) -> None:
        if chunk_games < 1:
            raise ValueError("chunk_games must be positive")
        self.chunk_games = int(chunk_games)
        self.profile = oracle_game_profile
        self.device = device
        self._waiting: dict[str, dict[str, Any]] = {str(b["block_id"]): dict(b) for b in blocks}
        self._answers: dict[tuple[str, int, int], dict[str, Any]] = {}
        self.launch_rounds = 0          # batched advances performed (for reports and tests)
        self.unforeseen_calls = 0

    def _bound(self, block: Mapping[str, Any]) -> int:
        """The stage's per-call attempt bound (``plan_h2h_chunk``, h2h_schedule.py:94-129)."""
        attempted = int(block.get("games_attempted", 0))
        return min(int(block["max_attempts"]), attempted + self.chunk_games) - attempted

    _IDENTITY = ("root_seed", "pair_id", "order", "seat1_strategy", "seat2_strategy",
                 "n_completed_required", "max_attempts")
    _PROGRESS = ("games_attempted", "games_completed", "games_safety_limit", "wins_seat1", "wins_seat2")

    def _state(self, block: Mapping[str, Any]) -> tuple:
        return (tuple(block.get(f) for f in self._IDENTITY),
                tuple(int(block.get(f, 0)) for f in self._PROGRESS))

    def _advance_all(self, manifest: Any) -> None:
        current = [b for b in self._waiting.values() if self._bound(b) > 0
                   and int(b.get("games_completed", 0)) < int(b["n_completed_required"])]
        self._waiting.clear()
        while current:
            results = simulate_blocks(current, manifest, self.chunk_games, self.profile, device=self.device)
            self.launch_rounds += 1
            nxt = []
            for before, after in zip(current, results):
                key = (str(before["block_id"]), int(before.get("games_attempted", 0)), self._bound(before))
                self._answers[key] = (self._state(before), after)
                if after["completion_status"] == PARTIAL_RESUMABLE:
                    nxt.append(after)
            current = nxt

    def __call__(self, block: dict[str, Any], strategy_manifest: Any, chunk_games: int) -> dict[str, Any]:
        manifest = _load_manifest(strategy_manifest) if isinstance(strategy_manifest, (str, Path)) \
            else strategy_manifest
        if self._waiting:
            self._advance_all(manifest)
        key = (str(block["block_id"]), int(block.get("games_attempted", 0)), int(chunk_games))
        hit = self._answers.pop(key, None)
        if hit is not None and hit[0] == self._state(block):
            return hit[1]
        self.unforeseen_calls += 1
        return _simulate_block_from_manifest(block, manifest, chunk_games, self.profile, device=self.device)


def build_strategy_manifest(strategies: Sequence[ThresholdStrategy]):
    """Manifest DataFrame mapping ids to attributes (simulation/strategies.py:725-748)."""
    import pandas as pd

    rows: dict[int, dict[str, Any]] = {}
    for s in strategies:
        if s.strategy_id is None or int(s.strategy_id) in rows:
            continue
        attrs = {f: getattr(s, f) for f in STRATEGY_TUPLE_FIELDS}
        attrs["favor_dice_or_score"] = attrs["favor_dice_or_score"].value
        attrs["strategy_id"] = int(s.strategy_id)
        attrs["strategy_str"] = str(s)
        rows[int(s.strategy_id)] = attrs
    manifest = pd.DataFrame(rows.values())
    if not manifest.empty:
        manifest["strategy_id"] = manifest["strategy_id"].astype("Int32")
        manifest = manifest.sort_values("strategy_id", kind="mergesort").reset_index(drop=True)
    return manifest


__all__ = ["BatchedBlockRunner", "PARTIAL_RESUMABLE", "build_strategy_manifest", "gpu_block_runner",
           "parse_strategy_identifier", "simulate_blocks", "_block_progress",
           "_simulate_block_from_manifest"]
