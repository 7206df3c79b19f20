#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 1500 $PYT tests/ > gpurun_out/all.log 2>&1; echo "all exit $?" > gpurun_out/info.log
for m in 1 0; do
AMOE_TC_STAGE_OUT=$m timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_so$m.log 2> gpurun_out/bench_so$m.err; echo "bench so$m exit $?" >> gpurun_out/info.log
done
cat gpurun_out/info.log; tail -6 gpurun_out/all.log
python - <<'PY'
import json
for n in ("so1","so0"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1]); print(n, round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["gpu_launches"])
    except Exception as e: print(n, "ERR", e)
PY
