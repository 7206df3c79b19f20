#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "tf32 or gate or policy" > gpurun_out/mlp.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/mlp.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gate_fused|policy_head" -s 2 -c 4 --csv --log-file gpurun_out/mlp_u.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[]
for r in csv.reader(open("gpurun_out/mlp_u.csv")):
    if len(r)>14 and r[0].isdigit() and r[12]=="gpu__time_duration.sum": t.append((r[4][:18], round(float(r[14])/1e3,1)))
print(t)
PY
