#!/bin/bash
# Round-2 visit N: full GPU suite, short bench (pipelined), per-cell kernel times.
TAG=${1:-r2n}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --ref-shuffles 0 --parquet-batches 0 --strong-reps 0 \
    > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('$OUT/bench_$TAG.json')); print(b['value'], b['ms_per_step'], b['e2e']['value'], b['roofline'].get('frac_model'), b['roofline'].get('frac_executed'), b['parity_check']['equal'])"
: > $OUT/cells_$TAG.log
for k in 2 4 6 12; do timeout 100 python scripts/profile_cell.py $k 4300 2 >> $OUT/cells_$TAG.log 2>&1; done; cat $OUT/cells_$TAG.log
