"""Coordinate-derived RNG identities (host side).

Mirror of the reference's ``farkle.utils.random`` (src/farkle/utils/random.py) for the
names the simulation path uses: ``RandomPurpose`` (:18-37), ``coordinate_entropy``
(:80-124), ``coordinate_seed`` (:191-225) and ``spawn_seeds`` (:275-295).

Streams themselves are never built on the host: every (game, seat) generator is seeded
inside the CUDA kernels from the same 18-word entropy (csrc/rng.cuh).  What lives here is
the scalar bookkeeping the reference's runner does in Python — the entropy tuple and the
diagnostic fingerprints that label shuffles and games — in exact integer arithmetic, plus
bulk fingerprints through the C ABI (``fb_coordinate_seeds``).
"""

from __future__ import annotations

from enum import IntEnum
from typing import Final

import numpy as np

RNG_SCHEME_VERSION: Final = 2
MAX_UINT32: Final = 2**32 - 1
MAX_UINT64: Final = 2**64 - 1


class RandomPurpose(IntEnum):
    """Permanent integer namespaces (utils/random.py:18-37)."""

    INDEXED_SEED = 1
    PLAYER = 10
    STRATEGY = 11
    TOURNAMENT_SHUFFLE = 100
    SHUFFLE_PERMUTATION = 101
    TOURNAMENT_GAME = 102
    TOURNAMENT_PLAYER = 103
    H2H_PAIR = 200
    H2H_ORDER = 201
    H2H_GAME = 202
    H2H_PLAYER = 203
    TRUESKILL_DIAGNOSTIC = 300
    BOOTSTRAP = 400
    ROOT_STABILITY_BOOTSTRAP = 401
    TIE_BREAK = 500
    HGB = 600
    SEED_SELECTION = 700


def _uint64_words(value: int, *, name: str) -> tuple[int, int]:
    if isinstance(value, bool) or not 0 <= int(value) <= MAX_UINT64:
        raise ValueError(f"{name} must be an integer in [0, 2**64 - 1]")
    v = int(value)
    return v & MAX_UINT32, v >> 32


def _resolve_alias(primary, alias, *, primary_name: str, alias_name: str) -> int:
    if primary is not None and alias is not None and int(primary) != int(alias):
        raise ValueError(f"{primary_name} and {alias_name} identify different coordinates")
    value = primary if primary is not None else alias
    return 0 if value is None else int(value)


def coordinate_entropy(purpose, *, root_seed: int, k: int = 0, shuffle_index: int = 0,
                       pair_index: int | None = None, pair_id: int | None = None, order: int = 0,
                       game_index: int | None = None, attempt_index: int | None = None,
                       seat_index: int = 0, replicate_index: int = 0) -> tuple[int, ...]:
    """The 18 uint32 words every stream is keyed by (utils/random.py:80-124)."""
    try:
        namespace = RandomPurpose(int(purpose))
    except (TypeError, ValueError) as exc:
        raise ValueError(f"unregistered RNG purpose namespace: {purpose!r}") from exc
    pair = _resolve_alias(pair_index, pair_id, primary_name="pair_index", alias_name="pair_id")
    game = _resolve_alias(game_index, attempt_index, primary_name="game_index",
                          alias_name="attempt_index")
    entropy: list[int] = [RNG_SCHEME_VERSION, int(namespace)]
    for name, value in (("root_seed", root_seed), ("k", k), ("shuffle_index", shuffle_index),
                        ("pair_id", pair), ("order", order), ("game_index", game),
                        ("seat_index", seat_index), ("replicate_index", replicate_index)):
        entropy.extend(_uint64_words(value, name=name))
    return tuple(entropy)


# SeedSequence constants (NumPy bit_generator.pyx; see csrc/rng.cuh for the device twin)
_INIT_A, _MULT_A = 0x43B0D7E5, 0x931E8875
_INIT_B, _MULT_B = 0x8B51F9DD, 0x58F38DED
_MIX_L, _MIX_R = 0xCA01F9DD, 0x4973F715
_M32 = 0xFFFFFFFF


def _seedseq_words(entropy: tuple[int, ...], n_words: int) -> list[int]:
    """``SeedSequence(entropy).generate_state(n_words, uint32)`` in exact Python integers."""
    hc = _INIT_A

    def hashmix(v: int) -> int:
        nonlocal hc
        v = (v ^ hc) & _M32
        hc = (hc * _MULT_A) & _M32
        v = (v * hc) & _M32
        return v ^ (v >> 16)

    def mix(x: int, y: int) -> int:
        r = (_MIX_L * x - _MIX_R * y) & _M32
        return r ^ (r >> 16)

    e = list(entropy) + [0] * max(0, 4 - len(entropy))
    pool = [hashmix(e[i]) for i in range(4)]
    for s in range(4):
        for d in range(4):
            if s != d:
                pool[d] = mix(pool[d], hashmix(pool[s]))
    for s in range(4, len(entropy)):
        for d in range(4):
            pool[d] = mix(pool[d], hashmix(e[s]))
    hc = _INIT_B
    out = []
    for i in range(n_words):
        v = pool[i & 3] ^ hc
        hc = (hc * _MULT_B) & _M32
        v = (v * hc) & _M32
        out.append(v ^ (v >> 16))
    return out


def coordinate_seed(purpose, *, dtype=np.uint64, **coords) -> int:
    """Diagnostic fingerprint of a coordinate (utils/random.py:191-225)."""
    w = _seedseq_words(coordinate_entropy(purpose, **coords), 2)
    if np.dtype(dtype) == np.dtype(np.uint32):
        return w[0]
    if np.dtype(dtype) == np.dtype(np.uint64):
        return w[0] | (w[1] << 32)
    raise ValueError("dtype must be numpy.uint32 or numpy.uint64")


def spawn_seeds(n: int, *, seed: int) -> np.ndarray:
    """Legacy external-boundary seeds (utils/random.py:275-295): INDEXED_SEED fingerprints."""
    if isinstance(n, bool) or n < 0:
        raise ValueError("n must be a non-negative integer")
    return np.asarray([coordinate_seed(RandomPurpose.INDEXED_SEED, root_seed=seed, game_index=i,
                                       dtype=np.uint32) for i in range(n)], dtype=np.uint32)


def coordinate_seeds_device(purpose, *, vary: str, base: int = 0, n: int, dtype=np.uint64,
                            device: int | None = None, **coords) -> np.ndarray:
    """Bulk fingerprints on the GPU: coordinate ``vary`` runs over ``base .. base+n-1``."""
    from .device import get_engine

    coordinate_entropy(purpose, **coords)  # validates the fixed coordinates
    out = get_engine(device).coordinate_seeds(
        int(purpose), root_seed=int(coords["root_seed"]), k=int(coords.get("k", 0)),
        shuffle_index=int(coords.get("shuffle_index", 0)),
        pair_id=int(coords.get("pair_id") or coords.get("pair_index") or 0),
        order=int(coords.get("order", 0)),
        game_index=int(coords.get("game_index") or coords.get("attempt_index") or 0),
        vary=vary, base=base, n=n, as_u32=np.dtype(dtype) == np.dtype(np.uint32))
    return out.astype(dtype)


__all__ = ["MAX_UINT32", "MAX_UINT64", "RNG_SCHEME_VERSION", "RandomPurpose", "coordinate_entropy",
           "coordinate_seed", "coordinate_seeds_device", "spawn_seeds"]
