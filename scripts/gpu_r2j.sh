#!/bin/bash
# Round-2 visit J (8 GPUs): scaling bench incl. the strong-scaling leg.  Bounded: 200 s per run.
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --ref-shuffles 0 --steps 10 --warmup 3 > $OUT/bench_r2j_$N.json 2> $OUT/bench_r2j_$N.err; echo "benchN rc=$?"; tail -2 $OUT/bench_r2j_$N.err
python -c "
import json; b=json.load(open('$OUT/bench_r2j_$N.json')); print('N=$N', b['value'], b['ms_per_step'], b['e2e']['value']); print(json.dumps(b['strong']))"
