#!/bin/bash
# GPU call 11: dual MMA issuers in the flat conv, CUDA-graph replay, chunk-size A/B under graphs; ncu --set full of stem_pool
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 120 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_automoe.py tests/test_gpu_kernels.py > gpurun_out/t_new.log 2>&1; echo "t_new exit $?" > gpurun_out/info.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
for c in 0 16 32 64; do
  AMOE_L2_CHUNK=$c timeout 300 $B > gpurun_out/bench_c$c.log 2> gpurun_out/bench_c$c.err; echo "bench graph chunk=$c exit $?" >> gpurun_out/info.log
done
AMOE_L2_CHUNK=0 timeout 300 $B --no-graph > gpurun_out/bench_eager_c0.log 2> gpurun_out/bench_eager_c0.err; echo "bench eager chunk=0 exit $?" >> gpurun_out/info.log
AMOE_L2_CHUNK=0 AMOE_OVERLAP=0 timeout 300 $B > gpurun_out/bench_c0_noov.log 2> gpurun_out/bench_c0_noov.err; echo "bench graph chunk=0 no-overlap exit $?" >> gpurun_out/info.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo "bench full(e2e) exit $?" >> gpurun_out/info.log
export AMOE_L2_CHUNK=0
ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel" -s 2 -c 1 \
    -o gpurun_out/prof4 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/info.log
ncu -i gpurun_out/prof4.ncu-rep --page raw --csv > gpurun_out/prof4_raw.csv 2> gpurun_out/raw.err
ncu -i gpurun_out/prof4.ncu-rep --page source --csv > gpurun_out/prof4_src.csv 2>/dev/null
cat gpurun_out/info.log; tail -3 gpurun_out/t_new.log; for f in gpurun_out/bench_*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3), (d.get("e2e") or {}).get("value"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
