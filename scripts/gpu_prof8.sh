#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
AMOE_FLAT_RES_PREFETCH=$m ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"flat" -s 7 -c 7 --csv --log-file gpurun_out/flat_rp$m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[]
for r in csv.reader(open("gpurun_out/flat_rp$m.csv")):
    if len(r)>14 and r[0].isdigit() and r[12]=="gpu__time_duration.sum": t.append(round(float(r[14])/1e3,1))
print("res_prefetch=$m", t)
PY
done
