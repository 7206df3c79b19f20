#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for m in 0 1 auto; do
if [ $m = auto ]; then unset AMOE_TC_EPI8; else export AMOE_TC_EPI8=$m; fi
timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_e$m.log 2> gpurun_out/bench_e$m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_e$m.log").read().strip().splitlines()[-1]); print("epi8=$m", round(d["value"]), d["ms_per_step"])
PY
done
done
