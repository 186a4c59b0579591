"""Where the unmodified reference checkout lives on this box.

In the build container it is /root/reference (read-only).  On the GPU box that path does not
exist; ``scripts/stage_reference.sh`` (run by ``__graft_entry__.build()``) stages a verbatim copy of
its importable tree under the git-ignored ``baseline/_ref/``, which travels with the repo snapshot.
The staged copy is preferred so that both boxes exercise the same files.
"""

from __future__ import annotations

import os
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def reference_root() -> Path | None:
    for cand in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if (cand / "src" / "farkle" / "simulation" / "run_tournament.py").is_file():
            return cand
    return None


def numba_cache_env() -> None:
    """The reference's @njit(cache=True) modules need a writable cache directory."""
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
