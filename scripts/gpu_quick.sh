#!/bin/bash
# Quick GPU visit: parity tests + per-cell kernel timings.   bash scripts/gpu_quick.sh TAG [extra cells...]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
: > $OUT/cells_$TAG.log
for spec in "2 4300" "4 4300" "$@"; do python scripts/profile_cell.py $spec 3 >> $OUT/cells_$TAG.log 2>&1; done
cat $OUT/cells_$TAG.log
# per-kernel times of one k=2 cell (ncu launch list: cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/cell_launches_$TAG.csv \
    python scripts/profile_cell.py 2 4300 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/cell_launches_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4].split("(")[0][-40:], float(r[-1])/1e6, "ms")
PY
