#!/usr/bin/env python
"""Build a variant of the library with extra -D defines: python scripts/build_variant.py NAME [-DFOO=1 ...]

Writes farkle_ii_b200/libfarkle_b200_NAME.so (git-ignored, travels with gpurun); run with
FARKLE_B200_LIB=farkle_ii_b200/libfarkle_b200_NAME.so.  Without defines NAME must be one of
`_native.TEST_VARIANTS` (e.g. `rejects`).
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from farkle_ii_b200 import _native  # noqa: E402

name, defines = sys.argv[1], sys.argv[2:]
print(_native.build_variant(name, defines or None, force=True))
