#!/bin/bash
# TF32 mma.sync gate / policy-head variants: kernel test, full suite, smoke, bench A/B
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "tf32" > gpurun_out/tf32.log 2>&1; echo "tf32 exit $?" > gpurun_out/info.log
timeout 1500 $PYT tests/ > gpurun_out/all.log 2>&1; echo "all exit $?" >> gpurun_out/info.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_tc1.log 2> gpurun_out/bench_tc1.err; echo "bench tc1 exit $?" >> gpurun_out/info.log
AMOE_MLP_TC=0 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_tc0.log 2> gpurun_out/bench_tc0.err; echo "bench tc0 exit $?" >> gpurun_out/info.log
cat gpurun_out/info.log; tail -15 gpurun_out/tf32.log; tail -8 gpurun_out/all.log; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
for n in ("tc1","tc0"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1]); print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["gpu_launches"])
    except Exception as e: print(n, "ERR", e)
PY
