#!/bin/bash
# Round-2 visit A: issue-peak probe variants (+ncu), ncu --set full of play_kernel for k = 4, 6, 12.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi_r2a.txt 2>&1
python scripts/issue_peak.py > $OUT/issue_peak_r2a.log 2>&1; cat $OUT/issue_peak_r2a.log
ncu --set full --clock-control none -k regex:issue_peak -c 10 -o $OUT/issue_peak_r2a -f \
    python scripts/issue_peak.py > $OUT/ncu_issue_peak_r2a.log 2>&1; echo "ncu probe rc=$?"
: > $OUT/cells_r2a.log
for k in 2 4 6 12; do python scripts/profile_cell.py $k 4300 2 >> $OUT/cells_r2a.log 2>&1; done
cat $OUT/cells_r2a.log
for k in 4 6 12; do
  ncu --set full --clock-control none --import-source on -k regex:play_kernel -s 1 -c 1 -o $OUT/play_k${k}_r2a -f \
      python scripts/profile_cell.py $k 4300 2 > $OUT/ncu_play_k${k}_r2a.log 2>&1; echo "ncu play k=$k rc=$?"
done
ls -la $OUT/*r2a*
