#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "tf32" > gpurun_out/tf32.log 2>&1; echo "tf32 exit $?" > gpurun_out/info.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_tc1.log 2> gpurun_out/bench_tc1.err; echo "bench tc1 exit $?" >> gpurun_out/info.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gate|policy_head|mean_hw" -c 6 --csv --log-file gpurun_out/launches_mlp.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
cat gpurun_out/info.log; tail -5 gpurun_out/tf32.log
grep -E "gate|policy|mean" gpurun_out/launches_mlp.csv | cut -d, -f5,15 | cut -c1-100
python - <<'PY'
import json
for n in ("tc1",):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1]); print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["gpu_launches"])
    except Exception as e: print(n, "ERR", e)
PY
