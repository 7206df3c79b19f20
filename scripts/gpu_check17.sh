#!/bin/bash
# GPU call 17: folded stem (scale in filters, bias on the padding channel): parity + A/B
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_kernels.py tests/test_gpu_automoe.py -s > gpurun_out/t_new.log 2>&1; echo "t_new exit $?" > gpurun_out/info.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/bench_fold1.log 2> gpurun_out/bench_fold1.err; echo "bench fold=1 exit $?" >> gpurun_out/info.log
AMOE_STEM_FOLD=0 timeout 300 $B > gpurun_out/bench_fold0.log 2> gpurun_out/bench_fold0.err; echo "bench fold=0 exit $?" >> gpurun_out/info.log
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:"stem_pool" -s 2 -c 2 --csv --log-file gpurun_out/launches_stem.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
cat gpurun_out/info.log; grep -E "passed|failed|bf16 rel err" gpurun_out/t_new.log | tail -8; for f in gpurun_out/bench_fold*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
grep "stem_pool" gpurun_out/launches_stem.csv | awk -F'","' '{print $13, $15}'
