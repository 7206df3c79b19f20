#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 1500 $PYT tests/ > gpurun_out/all.log 2>&1; echo "all exit $?" > gpurun_out/info.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?" >> gpurun_out/info.log
cat gpurun_out/info.log; tail -4 gpurun_out/all.log; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1]); print(round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["gpu_launches"])
PY
