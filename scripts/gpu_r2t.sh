#!/bin/bash
# Round-2 visit T: parity suite, smoke, bench (both arms), ncu launch list of the bench command.
TAG=${1:-r2t}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 $OUT/bench_$TAG.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
cut -c1-300 $OUT/bench_ref_$TAG.json; echo
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --cpu-seconds 1 --ref-shuffles 0 --parquet-batches 0 --strong-reps 0 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
