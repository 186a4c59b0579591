#!/bin/bash
# ncu --set full capture of play_kernel for one cell.   bash scripts/gpu_ncu.sh TAG K SHUFFLES
TAG=${1:-n}; K=${2:-2}; SH=${3:-4300}
OUT=gpurun_out; mkdir -p $OUT
python scripts/profile_cell.py $K $SH 2 > $OUT/cells_$TAG.log 2>&1; cat $OUT/cells_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:play_kernel -s 1 -c 1 -o $OUT/play_$TAG -f \
    python scripts/profile_cell.py $K $SH 2 > $OUT/ncu_play_$TAG.log 2>&1; echo "ncu play rc=$?"
