#!/bin/bash
# GPU call 12: stem epilogue v3 (FFMA2, ReLU after pooling, table-free horizontal pass); flat-conv tap-alignment timing experiment
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
export AMOE_L2_CHUNK=0
PYT="python -m pytest -m gpu -q --tb=short --timeout 120 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_automoe.py tests/test_gpu_kernels.py > gpurun_out/t_new.log 2>&1; echo "t_new exit $?" > gpurun_out/info.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/bench_base.log 2> gpurun_out/bench_base.err; echo "bench exit $?" >> gpurun_out/info.log
for m in 1 2; do
  AMOE_FLAT_DBG=$m timeout 300 $B > gpurun_out/bench_dbg$m.log 2> gpurun_out/bench_dbg$m.err; echo "bench dbg=$m exit $?" >> gpurun_out/info.log
done
for m in 0 1 2; do
AMOE_FLAT_DBG=$m ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:"conv3x3_flat|stem_pool" -s 16 -c 8 --csv --log-file gpurun_out/launches_dbg$m.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run$m.log 2>&1
done
cat gpurun_out/info.log; tail -3 gpurun_out/t_new.log; for f in gpurun_out/bench_*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
grep -h "gpu__time_duration\|tensor" gpurun_out/launches_dbg*.csv | awk -F'","' '{print FILENAME, $5, $13, $15}' | cut -c1-150
