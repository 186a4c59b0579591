#!/bin/bash
# Scaling visit on one 8-GPU box: bench.py at N = 2, 4, 8 the way the driver launches it.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale.sh TAG'
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
port=29511
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --steps 6 --warmup 3 > $OUT/bench_${n}gpu_$TAG.json 2> $OUT/bench_${n}gpu_$TAG.err
  echo "N=$n rc=$?"; cut -c1-260 $OUT/bench_${n}gpu_$TAG.json
  port=$((port + 1))
done
