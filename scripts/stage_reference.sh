#!/bin/bash
# Stage the UNMODIFIED reference checkout for the GPU box.
#
# /root/reference does not exist on the GPU box, but git-ignored files under the repo travel with the
# `gpurun` snapshot (like the built .so files).  This copies the reference's importable tree
# (src/, configs/, pyproject.toml -- nothing else) to baseline/_ref/, which is listed in .gitignore
# and NOT in .gpurunignore: it never enters this repo's history, and the GPU tests / bench.py can
# import the real reference there (drop-in proof on the CUDA engine, CPU baseline of the real
# Python path).  No file is edited.  Idempotent.
set -euo pipefail
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
[ -d "$SRC/src/farkle" ] || { echo "no reference checkout at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src" "$DST/src"
cp -r "$SRC/configs" "$DST/configs"
cp "$SRC/pyproject.toml" "$DST/pyproject.toml"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC" && find src configs pyproject.toml -type f ! -path '*/__pycache__/*' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "staged $(find "$DST/src" -name '*.py' | wc -l) python files -> $DST"
