#!/bin/bash
# Round-2 visit S: ncu --set full captures of play_kernel for k = 2, 4, 6, 12 (4,300-shuffle cells) and the
# step timelines.  Each capture only after the same command has exited 0 without ncu.
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/cells_r2s.log
for k in 2 4 6 12; do timeout 100 python scripts/profile_cell.py $k 4300 2 >> $OUT/cells_r2s.log 2>&1 || exit 1; done
cat $OUT/cells_r2s.log
for k in 2 4 6 12; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:play_kernel -s 1 -c 1 -o $OUT/play_r2s_k$k -f \
      python scripts/profile_cell.py $k 4300 2 > $OUT/ncu_play_r2s_k$k.log 2>&1; echo "ncu k=$k rc=$?"
done
timeout 200 python scripts/timeline_step.py 2 > $OUT/timeline_r2s_pipelined.txt 2>&1
timeout 200 python scripts/timeline_step.py --unpipelined 2 > $OUT/timeline_r2s_unpipelined.txt 2>&1
head -1 $OUT/timeline_r2s_pipelined.txt $OUT/timeline_r2s_unpipelined.txt
