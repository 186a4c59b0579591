#!/bin/bash
# Round-2 visit K: cell pipeline gated on the end of play_kernel.  Bounded runs.
TAG=${1:-r2k}
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q -k "play_cells or mega" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
for mode in "" "--unpipelined"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --ref-shuffles 0 --parquet-batches 0 --strong-reps 0 $mode \
      > $OUT/bench_$TAG$mode.json 2> $OUT/bench_$TAG$mode.err; echo "bench $mode rc=$?"
  python -c "
import json; b=json.load(open('$OUT/bench_$TAG$mode.json')); print('$mode', b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline'].get('kernel_ms'), b['parity_check']['equal'])"
done
