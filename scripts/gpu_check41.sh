#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 120 --timeout-method=thread"
timeout 300 $PYT tests/test_gpu_kernels.py -k "flat" > gpurun_out/flat.log 2>&1; echo "flat tests exit $?"; tail -2 gpurun_out/flat.log
timeout 300 $PYT tests/test_gpu_automoe.py > gpurun_out/model.log 2>&1; echo "model tests exit $?"; tail -2 gpurun_out/model.log
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"flat" -s 7 -c 7 --csv --log-file gpurun_out/flat8.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[];u=[]
for r in csv.reader(open("gpurun_out/flat8.csv")):
    if len(r)>14 and r[0].isdigit():
        (t if r[12]=="gpu__time_duration.sum" else u).append(round(float(r[14])/ (1e3 if r[12]=="gpu__time_duration.sum" else 1),1))
print("us", t, "sum", round(sum(t))); print("tensor%", u)
PY
for rep in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_f8.log 2> gpurun_out/bench_f8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_f8.log").read().strip().splitlines()[-1]); print("bench", round(d["value"]), d["ms_per_step"])
PY
done
