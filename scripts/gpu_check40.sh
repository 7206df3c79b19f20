#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
AMOE_FLAT_DBG=$m timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"flat" -s 7 -c 7 --csv --log-file gpurun_out/flat_dbg$m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[];u=[]
for r in csv.reader(open("gpurun_out/flat_dbg$m.csv")):
    if len(r)>14 and r[0].isdigit():
        (t if r[12]=="gpu__time_duration.sum" else u).append(round(float(r[14])/ (1e3 if r[12]=="gpu__time_duration.sum" else 1),1))
print("dbg=$m us", t); print("   tensor%", u)
PY
done
