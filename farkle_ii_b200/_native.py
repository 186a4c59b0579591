"""Loader for the C-ABI CUDA library (``include/farkle_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc
-shared``; no torch headers are involved).  There is no CPU fallback: if the
shared object is missing or no CUDA device can be bound, every entry point
raises.
"""

from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
# FARKLE_B200_LIB: load another build of the same sources (kernel experiments: scripts/build_variant.py)
DEFAULT_LIB_PATH = PKG_DIR / "libfarkle_b200.so"  # what build() writes
LIB_PATH = Path(os.environ.get("FARKLE_B200_LIB") or DEFAULT_LIB_PATH)  # what lib() loads
HEADER = PKG_DIR.parent / "include" / "farkle_b200.h"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]

# every symbol include/farkle_b200.h declares
SYMBOLS = (
    "fb_abi_version", "fb_last_error", "fb_init", "fb_device_info", "fb_row_stride",
    "fb_workspace_bytes", "fb_seedseq_generate", "fb_coordinate_seeds", "fb_seed_streams",
    "fb_roll_dice", "fb_default_score", "fb_permute_shuffles", "fb_play_tournament",
    "fb_play_tournament_seats", "fb_play_tournament_lags", "fb_cells_workspace_bytes",
    "fb_play_tournament_cells", "fb_matchup_scratch_bytes",
    "fb_play_h2h", "fb_h2h_resolve", "fb_play_games", "fb_run_tournament_host",
    "fb_measure_issue_peak", "fb_measure_issue_peak_variant", "fb_last_play_kernel_ms", "fb_play_kernel_ms_history",
    "fb_kernel_launch_count", "fb_timeline", "fb_timeline_dump",
)


class LagRequest(C.Structure):
    """``fb_lag_request_t`` (include/farkle_b200.h)."""

    _fields_ = [
        ("lags", C.c_void_p), ("n_lags", C.c_int32),
        ("matchup_min_observations", C.c_int32),
        ("strategy_stats_dev", C.c_void_p), ("strategy_edges_dev", C.c_void_p),
        ("matchup_capacity", C.c_uint64), ("matchup_participants_dev", C.c_void_p),
        ("matchup_count_dev", C.c_void_p), ("matchup_stats_dev", C.c_void_p),
        ("scratch_dev", C.c_void_p), ("scratch_bytes", C.c_size_t),
        ("n_matchups_host", C.c_void_p), ("first_seen_dev", C.c_void_p),
        ("all_player_dev", C.c_void_p),
    ]


class Cell(C.Structure):
    """``fb_cell_t`` (include/farkle_b200.h)."""

    _fields_ = [("root_seed", C.c_uint64), ("shuffle0", C.c_uint64), ("k", C.c_int32),
                ("n_shuffles", C.c_int32), ("tallies_dev", C.c_void_p), ("totals_dev", C.c_void_p)]


class NativeError(RuntimeError):
    """A C-ABI call failed; the message is ``fb_last_error()``."""


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HEADER]


def is_stale() -> bool:
    if not DEFAULT_LIB_PATH.exists():
        return True
    built = DEFAULT_LIB_PATH.stat().st_mtime
    return any(src.stat().st_mtime > built for src in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile ``csrc/capi.cu`` for sm_100a with nvcc (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return DEFAULT_LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(DEFAULT_LIB_PATH), str(CSRC / "capi.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr)
    return DEFAULT_LIB_PATH


# Test builds of the same sources (never loaded by the product path; FARKLE_B200_LIB selects one):
#   rejects  Lemire threshold 2^30 instead of 4: one half in four is rejected, so the face queue's
#            handling of rejected halves runs all the time (tests/rejects_variant_check.py; the oracle
#            has the matching run-time knob FB_TEST_LEMIRE_THR)
TEST_VARIANTS = {"rejects": ["-DFB_LEMIRE_THR=0x40000000u"]}


def variant_path(name: str) -> Path:
    return PKG_DIR / f"libfarkle_b200_{name}.so"


def build_variant(name: str, defines: list[str] | None = None, force: bool = False) -> Path:
    """Compile ``csrc/capi.cu`` with extra ``-D`` defines into ``libfarkle_b200_<name>.so`` (in-tree,
    git-ignored, travels with ``gpurun``).  Up-to-date files are kept."""
    out = variant_path(name)
    defines = list(TEST_VARIANTS[name] if defines is None else defines)
    if not force and out.exists() and all(src.stat().st_mtime <= out.stat().st_mtime for src in sources()):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    proc = subprocess.run([nvcc, *NVCC_FLAGS, *defines, "-o", str(out), str(CSRC / "capi.cu")],
                          capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{proc.stdout}\n{proc.stderr}")
    return out


_lib: C.CDLL | None = None

_u64, _u32, _i32, _int, _vp, _sz = (C.c_uint64, C.c_uint32, C.c_int32, C.c_int, C.c_void_p,
                                    C.c_size_t)


def _declare(L: C.CDLL) -> None:
    L.fb_abi_version.restype = _int
    L.fb_last_error.restype = C.c_char_p
    L.fb_init.argtypes = [_int]
    L.fb_device_info.argtypes = [C.POINTER(_int)] * 4
    L.fb_row_stride.restype = _sz
    L.fb_row_stride.argtypes = [_int]
    L.fb_workspace_bytes.restype = _sz
    L.fb_workspace_bytes.argtypes = [_int, _u64]
    L.fb_seedseq_generate.argtypes = [_vp, _int, _u64, _int, _vp, _vp]
    L.fb_coordinate_seeds.argtypes = [_u32, _u64, _u64, _u64, _u64, _u64, _u64, _int, _u64, _u64,
                                      _int, _vp, _vp]
    L.fb_seed_streams.argtypes = [_vp, _u64, _vp, _vp]
    L.fb_roll_dice.argtypes = [_vp, _vp, _u64, _vp, _int, _vp, _vp]
    L.fb_default_score.argtypes = [_vp, _vp, _vp, _u64, _vp, _vp]
    L.fb_permute_shuffles.argtypes = [_u64, _int, _u64, _int, _int, _vp, _vp]
    L.fb_play_tournament.argtypes = [_u64, _int, _u64, _int, _vp, _vp, _int, _int, _i32, _i32,
                                     _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _int, _vp, _sz, _vp]
    L.fb_play_tournament_seats.argtypes = [_u64, _int, _u64, _int, _vp, _vp, _int, _int, _i32, _i32,
                                           _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _int, _vp, _vp,
                                           _sz, _vp]
    L.fb_play_tournament_lags.argtypes = [_u64, _int, _u64, _int, _vp, _vp, _int, _int, _i32, _i32,
                                          _vp, _vp, _vp, _int, _int, _vp, _vp, _vp, _int, _vp,
                                          C.POINTER(LagRequest), _vp, _sz, _vp]
    L.fb_cells_workspace_bytes.argtypes = [C.POINTER(Cell), _int, _int]
    L.fb_cells_workspace_bytes.restype = _sz
    L.fb_play_tournament_cells.argtypes = [C.POINTER(Cell), _int, _int, _vp, _vp, _int, _int, _i32, _i32,
                                           _int, _vp, _sz, _vp]
    L.fb_matchup_scratch_bytes.argtypes = [_u64]
    L.fb_matchup_scratch_bytes.restype = _sz
    L.fb_play_h2h.argtypes = [_u64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _u64, _i32, _i32, _vp,
                              _vp, _vp, _vp, _sz, _vp]
    L.fb_h2h_resolve.argtypes = [_int, _vp, _vp, _vp, _vp, _vp]
    L.fb_play_games.argtypes = [_vp, _u64, _int, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp,
                                _sz, _vp]
    L.fb_run_tournament_host.argtypes = [_u64, _int, _u64, _int, _vp, _vp, _int, _int, _i32, _i32,
                                         _int, _vp, _vp, _vp, _int]
    L.fb_measure_issue_peak.argtypes = [_int, C.POINTER(C.c_double)]
    L.fb_measure_issue_peak_variant.argtypes = [_int, _int, C.POINTER(C.c_double)]
    L.fb_last_play_kernel_ms.restype = C.c_float
    L.fb_play_kernel_ms_history.argtypes = [C.POINTER(C.c_float), _int]
    L.fb_kernel_launch_count.restype = _u64
    L.fb_timeline.argtypes = [_int]
    L.fb_timeline_dump.argtypes = [C.c_char_p, _sz]


def lib() -> C.CDLL:
    """Return the loaded library; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`."
                " farkle_ii_b200 has no CPU fallback.")
        L = C.CDLL(str(LIB_PATH))
        _declare(L)
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise NativeError(f"farkle_b200 error {rc}: {lib().fb_last_error().decode()}")
