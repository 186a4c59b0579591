#!/bin/bash
# Round-2 visit G (N GPUs): cells test, 1-GPU bench with the strong leg, then torchrun bench at N.
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q -x -k "play_cells" > $OUT/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_r2g.log
python bench.py --ref-shuffles 0 --cpu-seconds 2 > $OUT/bench_r2g_1.json 2> $OUT/bench_r2g_1.err; echo "bench1 rc=$?"; tail -2 $OUT/bench_r2g_1.err
python -c "
import json; b=json.load(open('$OUT/bench_r2g_1.json')); print('N=1', b['value'], b['ms_per_step'], json.dumps(b['strong'])[:600])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --ref-shuffles 0 > $OUT/bench_r2g_$N.json 2> $OUT/bench_r2g_$N.err; echo "benchN rc=$?"; tail -2 $OUT/bench_r2g_$N.err
python -c "
import json; b=json.load(open('$OUT/bench_r2g_$N.json')); print('N=$N', b['value'], b['ms_per_step'], json.dumps(b['strong'])[:900])"
