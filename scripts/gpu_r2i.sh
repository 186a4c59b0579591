#!/bin/bash
# Round-2 visit I: compact score table + K2 home slots: parity suite, cell timings, rows-to-Parquet leg, H2H stage timing.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_host_surface.py tests/test_rng_diagnostics.py -m gpu -q -x > $OUT/pytest_r2i.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_r2i.log
: > $OUT/cells_r2i.log
for k in 2 4 6 12; do python scripts/profile_cell.py $k 4300 3 >> $OUT/cells_r2i.log 2>&1; done
cat $OUT/cells_r2i.log
timeout 600 python bench.py --ref-shuffles 0 --cpu-seconds 2 --strong-reps 0 > $OUT/bench_r2i.json 2> $OUT/bench_r2i.err; echo "bench rc=$?"; tail -3 $OUT/bench_r2i.err
python -c "
import json; b=json.load(open('$OUT/bench_r2i.json'))
print('value', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'], b['roofline']['kernel_ms_by_k'])
print('e2e_parquet', json.dumps(b['e2e_parquet']))"
timeout 900 python scripts/h2h_stage_time.py 8 > $OUT/h2h_stage_r2i.json 2> $OUT/h2h_stage_r2i.err; echo "h2h stage rc=$?"; cut -c1-1500 $OUT/h2h_stage_r2i.json; tail -3 $OUT/h2h_stage_r2i.err
