#!/bin/bash
# Quick GPU visit: parity tests + per-cell kernel timings.   bash scripts/gpu_quick.sh TAG [extra cells...]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
: > $OUT/cells_$TAG.log
for spec in "2 4300" "4 4300" "$@"; do python scripts/profile_cell.py $spec 3 >> $OUT/cells_$TAG.log 2>&1; done
cat $OUT/cells_$TAG.log
