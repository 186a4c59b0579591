#!/bin/bash
# Round-2 visit E: cells pipeline: new tests, bench pipelined vs unpipelined, launch timeline.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_r2e.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_r2e.log
python bench.py --ref-shuffles 0 --cpu-seconds 2 > $OUT/bench_r2e.json 2> $OUT/bench_r2e.err; echo "bench rc=$?"; tail -3 $OUT/bench_r2e.err
python -c "
import json; b=json.load(open('$OUT/bench_r2e.json')); print('pipelined', b['value'], b['ms_per_step'], b['roofline']['kernel_ms_by_k'], b['gpu_launches'], b['parity_check']['equal'])"
python bench.py --ref-shuffles 0 --cpu-seconds 2 --unpipelined > $OUT/bench_r2e_unp.json 2> $OUT/bench_r2e_unp.err; echo "bench rc=$?"
python -c "
import json; b=json.load(open('$OUT/bench_r2e_unp.json')); print('unpipelined', b['value'], b['ms_per_step'], b['roofline']['kernel_ms_by_k'], b['gpu_launches'])"
for w in 24 28; do FB_PLAY_WARPS=$w python scripts/profile_cell.py 2 4300 2; FB_PLAY_WARPS=$w python scripts/profile_cell.py 4 4300 2; done
