#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
AMOE_MLP_PREFETCH=$m ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gate_fused|policy_head" -s 2 -c 4 --csv --log-file gpurun_out/mlp_pf$m.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[]
for r in csv.reader(open("gpurun_out/mlp_pf$m.csv")):
    if len(r)>14 and r[0].isdigit() and r[12]=="gpu__time_duration.sum": t.append(round(float(r[14])/1e3,1))
print("mlp_prefetch=$m", t)
PY
done
for rep in 1 2; do for m in 0 1; do
AMOE_MLP_PREFETCH=$m timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_pf$m.log 2> gpurun_out/bench_pf$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pf$m.log").read().strip().splitlines()[-1]); print("mlp_prefetch=$m", round(d["value"]), d["ms_per_step"])
PY
done; done
