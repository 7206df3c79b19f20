#!/bin/bash
mkdir -p gpurun_out
for d in 4 8 16 28 6; do
AMOE_KW3_DBG=$d timeout 600 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_dbg$d.log 2> gpurun_out/bench_dbg$d.err; echo "dbg$d exit $?"
done
python - <<'PY'
import json
for n in ("dbg4","dbg8","dbg16","dbg28","dbg6"):
    try:
        d=json.loads(open(f"gpurun_out/bench_{n}.log").read().strip().splitlines()[-1]); print(n, round(d["value"]), d["ms_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
