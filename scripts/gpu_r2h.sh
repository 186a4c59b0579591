#!/bin/bash
# Round-2 visit H: whole GPU suite, bench (all legs), H2H stage timing.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_r2h.log 2>&1; echo "pytest rc=$?"; tail -14 $OUT/pytest_r2h.log
timeout 600 python bench.py > $OUT/bench_r2h.json 2> $OUT/bench_r2h.err; echo "bench rc=$?"; tail -3 $OUT/bench_r2h.err
python -c "
import json; b=json.load(open('$OUT/bench_r2h.json'))
print('value', b['value'], 'ms', b['ms_per_step'], 'e2e', b['e2e']['value'])
print('e2e_parquet', json.dumps(b['e2e_parquet']))
print('strong', b['strong']['n1_ms']); print('pyref', json.dumps(b['cpu_baseline_reference'])[:300]); print('parity', b['parity_check']['equal'])"
timeout 900 python scripts/h2h_stage_time.py 8 > $OUT/h2h_stage_r2h.json 2> $OUT/h2h_stage_r2h.err; echo "h2h stage rc=$?"; cat $OUT/h2h_stage_r2h.json; tail -3 $OUT/h2h_stage_r2h.err
