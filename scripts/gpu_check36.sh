#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_kernels.py -k "upsample" tests/test_gpu_automoe.py > gpurun_out/up.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/up.log
ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:"upsample" -s 3 -c 3 --csv --log-file gpurun_out/up.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
for r in csv.reader(open("gpurun_out/up.csv")):
    if len(r)>14 and r[0].isdigit(): print(r[4][:40], r[8], r[12], r[14])
PY
for rep in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_up.log 2> gpurun_out/bench_up.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_up.log").read().strip().splitlines()[-1]); print("bench", round(d["value"]), d["ms_per_step"])
PY
done
