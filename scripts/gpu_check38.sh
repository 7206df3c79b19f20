#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 120 --timeout-method=thread -x"
timeout 400 $PYT tests/test_gpu_kernels.py -k "conv" > gpurun_out/conv.log 2>&1; echo "conv tests exit $?"; tail -3 gpurun_out/conv.log
timeout 400 $PYT tests/test_gpu_automoe.py > gpurun_out/model.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/model.log
for m in 0 1; do
AMOE_TC_M2=$m timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"conv_tc_kernel" -s 20 -c 20 --csv --log-file gpurun_out/tc_m2_$m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
t=[];u=[]
for r in csv.reader(open("gpurun_out/tc_m2_$m.csv")):
    if len(r)>14 and r[0].isdigit():
        if r[12]=="gpu__time_duration.sum": t.append(round(float(r[14])/1e3))
        else: u.append(round(float(r[14])))
print("m2=$m us", t, "sum", sum(t)); print("      tensor%", u)
PY
done
for rep in 1 2; do for m in 0 1; do
AMOE_TC_M2=$m timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_m2$m.log 2> gpurun_out/bench_m2$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_m2$m.log").read().strip().splitlines()[-1]); print("m2=$m", round(d["value"]), d["ms_per_step"])
PY
done; done
