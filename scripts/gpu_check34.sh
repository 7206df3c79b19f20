#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "flat" > gpurun_out/flat.log 2>&1; echo "flat exit $?"; tail -3 gpurun_out/flat.log
for rep in 1 2; do
for m in 0 1; do
AMOE_FLAT_RES_PREFETCH=$m timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_rp$m.log 2> gpurun_out/bench_rp$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rp$m.log").read().strip().splitlines()[-1]); print("res_prefetch=$m", round(d["value"]), d["ms_per_step"])
PY
done
done
