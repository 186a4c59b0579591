#!/bin/bash
# Round-2 visit U: ncu --set full captures of the final play_kernel (k = 2 and 4) for the SASS model.
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/cells_r2u.log
for k in 2 4; do timeout 100 python scripts/profile_cell.py $k 4300 2 >> $OUT/cells_r2u.log 2>&1 || exit 1; done
cat $OUT/cells_r2u.log | cut -c1-150
for k in 2 4; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:play_kernel -s 1 -c 1 -o $OUT/play_r2u_k$k -f \
      python scripts/profile_cell.py $k 4300 2 > $OUT/ncu_play_r2u_k$k.log 2>&1; echo "ncu k=$k rc=$?"
done
