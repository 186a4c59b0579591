#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu --set full of play_kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cut -c1-600 $OUT/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
cut -c1-400 $OUT/bench_ref_$TAG.json
for k in 2 4 6 12; do python scripts/profile_cell.py $k 4300 2 >> $OUT/cells_$TAG.log 2>&1; done
cat $OUT/cells_$TAG.log
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
# full capture of the dominant kernel (k=2 cell at bench size)
ncu --set full --clock-control none --import-source on -k regex:play_kernel -s 1 -c 1 -o $OUT/play_$TAG -f \
    python scripts/profile_cell.py 2 4300 2 > $OUT/ncu_play_$TAG.log 2>&1; echo "ncu play rc=$?"
ncu --set full --clock-control none --import-source on -k regex:finish_kernel -s 1 -c 1 -o $OUT/finish_$TAG -f \
    python scripts/profile_cell.py 2 4300 2 > $OUT/ncu_finish_$TAG.log 2>&1; echo "ncu finish rc=$?"
