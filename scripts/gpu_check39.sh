#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
AMOE_TC_STAGE_OUT=$m timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_tc_kernel" -s 20 -c 20 --csv --log-file gpurun_out/tc_e_$m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[]
for r in csv.reader(open("gpurun_out/tc_e_$m.csv")):
    if len(r)>14 and r[0].isdigit() and r[12]=="gpu__time_duration.sum": t.append(round(float(r[14])/1e3,1))
print("stage_out=$m us", t, "sum", round(sum(t)))
PY
done
