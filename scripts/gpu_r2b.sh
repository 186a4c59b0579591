#!/bin/bash
# Round-2 visit B: the whole GPU test suite (incl. the reference drop-in on the CUDA engine), smoke, bench both arms.
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_r2b.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_r2b.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_r2b.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_r2b.log
python bench.py > $OUT/bench_r2b.json 2> $OUT/bench_r2b.err; echo "bench rc=$?"; tail -3 $OUT/bench_r2b.err
cut -c1-1500 $OUT/bench_r2b.json
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_r2b.json 2> $OUT/bench_ref_r2b.err; echo "ref rc=$?"
cut -c1-600 $OUT/bench_ref_r2b.json
nproc; free -g | head -2
