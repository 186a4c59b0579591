"""ctypes front end of the CPU parity oracle (``oracle/farkle_oracle.c``).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; ``farkle_ii_b200`` never does.  Parity status: pinned against the
reference by ``tests/golden/make_golden.py`` -> ``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libfarkle_oracle.so"

TALLY_WIDTH = 26
TOTALS_WIDTH = 20
N_METRICS = 11

STRATEGY_DTYPE = np.dtype(
    [("score_threshold", "<i4"), ("dice_threshold", "<i2"), ("flags", "<u2")]
)

SF_SMART_FIVE = 0x01
SF_SMART_ONE = 0x02
SF_CONSIDER_SCORE = 0x04
SF_CONSIDER_DICE = 0x08
SF_REQUIRE_BOTH = 0x10
SF_AUTO_HOT_DICE = 0x20
SF_RUN_UP_SCORE = 0x40
SF_FAVOR_SCORE = 0x80

_SEAT_FIELDS = [
    ("score", "<i4"),
    ("strategy", "<i4"),
    ("highest_turn", "<i4"),
    ("farkles", "<u2"),
    ("rolls", "<u2"),
    ("n_turns", "<u2"),
    ("hot_dice", "<u2"),
    ("smart_five_uses", "<u2"),
    ("n_smart_five_dice", "<u2"),
    ("smart_one_uses", "<u2"),
    ("n_smart_one_dice", "<u2"),
]
SEAT_DTYPE = np.dtype(_SEAT_FIELDS)
assert SEAT_DTYPE.itemsize == 28


def row_stride(k: int) -> int:
    return (16 + 28 * k + 15) & ~15


def row_dtype(k: int) -> np.dtype:
    """Structured dtype matching ``fb_row_header_t`` + k x ``fb_row_seat_t``."""
    return np.dtype(
        {
            "names": ["game_seed", "game_ordinal", "n_rounds", "winner_seat", "flags", "seats"],
            "formats": ["<u8", "<u4", "<u2", "u1", "u1", (SEAT_DTYPE, (k,))],
            "offsets": [0, 8, 12, 14, 15, 16],
            "itemsize": row_stride(k),
        }
    )


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc if the shared object is missing or stale."""
    src = _HERE / "farkle_oracle.c"
    hdr = _HERE.parent / "include" / "farkle_b200.h"
    stale = (
        force
        or not _LIB_PATH.exists()
        or (src.exists() and _LIB_PATH.stat().st_mtime < src.stat().st_mtime)
        or (hdr.exists() and _LIB_PATH.stat().st_mtime < hdr.stat().st_mtime)
    )
    if stale:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


_lib: C.CDLL | None = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.fo_coordinate_seed.restype = C.c_uint64
        L.fo_row_stride.restype = C.c_size_t
        _lib = L
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _coords9(purpose, root_seed, k=0, shuffle_index=0, pair_id=0, order=0, game_index=0,
             seat_index=0, replicate_index=0) -> np.ndarray:
    return np.array(
        [purpose, root_seed, k, shuffle_index, pair_id, order, game_index, seat_index,
         replicate_index],
        dtype=np.uint64,
    )


def seedseq_generate(entropy, n_words: int) -> np.ndarray:
    e = np.ascontiguousarray(entropy, dtype=np.uint32)
    out = np.zeros(n_words, dtype=np.uint32)
    lib().fo_seedseq_generate(_p(e), C.c_int(len(e)), C.c_int(n_words), _p(out))
    return out


def coordinate_entropy(purpose, **kw) -> np.ndarray:
    c = _coords9(purpose, **kw)
    out = np.zeros(18, dtype=np.uint32)
    lib().fo_coordinate_entropy(_p(c), _p(out))
    return out


def coordinate_seed(purpose, *, as_u32: bool = False, **kw) -> int:
    c = _coords9(purpose, **kw)
    return int(lib().fo_coordinate_seed(_p(c), C.c_int(1 if as_u32 else 0)))


def seed_stream(purpose, **kw) -> np.ndarray:
    """Return ``[state_hi, state_lo, inc_hi, inc_lo]`` of ``coordinate_rng``."""
    c = _coords9(purpose, **kw)
    out = np.zeros(4, dtype=np.uint64)
    lib().fo_seed_stream(_p(c), _p(out))
    return out


def roll_dice_state(state_inc, n_dice, has32: int = 0, saved: int = 0) -> np.ndarray:
    si = np.ascontiguousarray(state_inc, dtype=np.uint64)
    nd = np.ascontiguousarray(n_dice, dtype=np.int32)
    out = np.zeros((len(nd), 6), dtype=np.uint8)
    lib().fo_roll_dice_state(_p(si), C.c_int(has32), C.c_uint32(saved), _p(nd), C.c_int(len(nd)),
                             _p(out))
    return out


def permutation(root_seed: int, k: int, shuffle_index: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    lib().fo_permutation(C.c_uint64(root_seed), C.c_uint64(k), C.c_uint64(shuffle_index),
                         C.c_int(n), _p(out))
    return out


def evaluate_counts(counts) -> tuple[int, int, int, int]:
    c = np.ascontiguousarray(counts, dtype=np.int32)
    out = np.zeros(4, dtype=np.int32)
    lib().fo_evaluate_counts(_p(c), _p(out))
    return tuple(int(x) for x in out)


def default_score(faces, turn_score_pre: int, strategy: np.ndarray) -> tuple[int, ...]:
    f = np.zeros(6, dtype=np.uint8)
    f[: len(faces)] = faces
    s = np.ascontiguousarray(strategy, dtype=STRATEGY_DTYPE).reshape(1)
    out = np.zeros(5, dtype=np.int32)
    lib().fo_default_score(_p(f), C.c_int32(turn_score_pre), _p(s), _p(out))
    return tuple(int(x) for x in out)


def play_tournament(root_seed, k, shuffle0, n_shuffles, strategies, *, strategy_ids=None,
                    n_tally_ids=None, target_score=10_000, max_rounds=200, overrides=(),
                    shuffles_per_slot=0, want_rows=False, want_game_seeds=False, n_threads=1):
    """Return ``(tallies[slots, ids, 26], totals[20], rows | None)``."""
    st = np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE)
    n = len(st)
    ids = None if strategy_ids is None else np.ascontiguousarray(strategy_ids, dtype=np.int32)
    if n_tally_ids is None:
        n_tally_ids = n if ids is None else int(ids.max()) + 1
    n_slots = 1 if shuffles_per_slot <= 0 else -(-n_shuffles // shuffles_per_slot)
    tallies = np.zeros((n_slots, n_tally_ids, TALLY_WIDTH), dtype=np.int64)
    totals = np.zeros(TOTALS_WIDTH, dtype=np.int64)
    gps = n // k
    rows = np.zeros(n_shuffles * gps, dtype=row_dtype(k)) if want_rows else None
    ov_s = np.array([o[0] for o in overrides], dtype=np.uint64)
    ov_g = np.array([o[1] for o in overrides], dtype=np.uint32)
    ov_m = np.array([o[2] for o in overrides], dtype=np.int32)
    rc = lib().fo_play_tournament(
        C.c_uint64(root_seed), C.c_int(k), C.c_uint64(shuffle0), C.c_int(n_shuffles), _p(st),
        _p(ids), C.c_int(n), C.c_int(n_tally_ids), C.c_int32(target_score), C.c_int32(max_rounds),
        _p(ov_s), _p(ov_g), _p(ov_m), C.c_int(len(overrides)), C.c_int(shuffles_per_slot),
        _p(tallies), _p(totals), _p(rows), C.c_int(int(want_game_seeds)), C.c_int(n_threads))
    if rc != 0:
        raise ValueError(f"fo_play_tournament failed: {rc}")
    return tallies, totals, rows


def play_games(coords, k, seat_strategies, *, seat_strategy_ids=None, target_score=10_000,
               max_rounds=200, target_scores=None, max_rounds_v=None):
    """Play games at explicit coordinates; return ``(rows, totals)``."""
    cc = np.ascontiguousarray(coords, dtype=np.uint64).reshape(-1, 7)
    n = len(cc)
    st = np.ascontiguousarray(seat_strategies, dtype=STRATEGY_DTYPE).reshape(n, k)
    ids = None if seat_strategy_ids is None else np.ascontiguousarray(
        seat_strategy_ids, dtype=np.int32).reshape(n, k)
    ts = None if target_scores is None else np.ascontiguousarray(target_scores, dtype=np.int32)
    mr = None if max_rounds_v is None else np.ascontiguousarray(max_rounds_v, dtype=np.int32)
    rows = np.zeros(n, dtype=row_dtype(k))
    totals = np.zeros(TOTALS_WIDTH, dtype=np.int64)
    rc = lib().fo_play_games(_p(cc), C.c_uint64(n), C.c_int(k), _p(st), _p(ids), _p(ts),
                             C.c_int32(target_score), _p(mr), C.c_int32(max_rounds), _p(rows),
                             _p(totals))
    if rc != 0:
        raise ValueError(f"fo_play_games failed: {rc}")
    return rows, totals


def play_h2h_block(root_seed, pair_id, order, seat1, seat2, *, n_completed_required,
                   max_attempts, chunk_games, progress=(0, 0, 0, 0, 0), target_score=10_000,
                   max_rounds=200):
    """Advance one H2H block; return ``(progress[5], outcomes)``."""
    s1 = np.ascontiguousarray(seat1, dtype=STRATEGY_DTYPE).reshape(1)
    s2 = np.ascontiguousarray(seat2, dtype=STRATEGY_DTYPE).reshape(1)
    pr = np.array(progress, dtype=np.int32)
    start = int(pr[0])
    oc = np.full(max(chunk_games, 1), 255, dtype=np.uint8)
    lib().fo_play_h2h_block(C.c_uint64(root_seed), C.c_uint64(pair_id), C.c_int(order), _p(s1),
                            _p(s2), C.c_int32(n_completed_required), C.c_int32(max_attempts),
                            C.c_int32(chunk_games), C.c_int32(target_score),
                            C.c_int32(max_rounds), _p(pr), _p(oc))
    return pr, oc[: int(pr[0]) - start]


def scan_rejected_halves(root_seed, k, shuffle0, n_shuffles, strategies, *, target_score=10_000,
                         max_rounds=200, cap=256) -> np.ndarray:
    """Games of the cell that meet a rejected Lemire half: rows ``(shuffle_index, game_index, rejected
    halves)`` (fixture helper for the rejection paths of the CUDA kernel; single-threaded)."""
    st = np.ascontiguousarray(strategies, dtype=STRATEGY_DTYPE)
    out = np.zeros((cap, 3), dtype=np.uint64)
    fn = lib().fo_scan_rejected_halves
    fn.restype = C.c_int64
    n = fn(C.c_uint64(root_seed), C.c_int(k), C.c_uint64(shuffle0), C.c_int(n_shuffles), _p(st),
           C.c_int(len(st)), C.c_int32(target_score), C.c_int32(max_rounds), _p(out), C.c_int64(cap))
    if n < 0:
        raise ValueError(f"fo_scan_rejected_halves failed: {n}")
    return out[: min(int(n), cap)]

