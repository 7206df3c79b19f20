#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "tf32 or gate or policy or cluster" > gpurun_out/tf32.log 2>&1; echo "tf32 exit $?" > gpurun_out/info.log
for m in 0 1; do
AMOE_MLP_TC=$m timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_tc$m.log 2> gpurun_out/bench_tc$m.err; echo "bench tc$m exit $?" >> gpurun_out/info.log
done
cat gpurun_out/info.log; tail -5 gpurun_out/tf32.log
python - <<'PY'
import json
for n in ("tc0","tc1"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1]); print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["gpu_launches"])
    except Exception as e: print(n, "ERR", e)
PY
