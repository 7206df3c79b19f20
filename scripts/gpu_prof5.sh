#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.per_cycle_active --clock-control none -k regex:"gate|policy_head|mean_hw" -c 3 --csv --log-file gpurun_out/launches_mlp.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/launches_mlp.csv')):
    if len(r)>14 and r[0].isdigit(): print(r[4][:40], r[8], r[12], r[14])
PY
