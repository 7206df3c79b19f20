#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
AMOE_STEM_DBG=$m timeout 120 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"stem_pool" -s 1 -c 1 --csv --log-file gpurun_out/stem_dbg$m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
for r in csv.reader(open("gpurun_out/stem_dbg$m.csv")):
    if len(r)>14 and r[0].isdigit(): print("dbg=$m", r[12][:40], r[14])
PY
done
