#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 900 $PYT tests/ > gpurun_out/all.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head1x1" -s 3 -c 3 --csv --log-file gpurun_out/h1.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<PY
import csv
for r in csv.reader(open("gpurun_out/h1.csv")):
    if len(r)>14 and r[0].isdigit(): print(r[4][:40], r[8], r[12], r[14])
PY
for rep in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_h1.log 2> gpurun_out/bench_h1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_h1.log").read().strip().splitlines()[-1]); print("bench", round(d["value"]), d["ms_per_step"])
PY
done
