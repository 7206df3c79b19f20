#!/bin/bash
# GPU call 21: cluster (DSMEM row-split) gate / policy-head kernels
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "gate or policy or mlp or submodules" > gpurun_out/t_mlp.log 2>&1; echo "t_mlp exit $?" > gpurun_out/info.log
timeout 900 $PYT tests/test_gpu_automoe.py > gpurun_out/t_model.log 2>&1; echo "t_model exit $?" >> gpurun_out/info.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/bench_cl1.log 2> gpurun_out/bench_cl1.err; echo "bench cluster=1 exit $?" >> gpurun_out/info.log
AMOE_MLP_CLUSTER=0 timeout 300 $B > gpurun_out/bench_cl0.log 2> gpurun_out/bench_cl0.err; echo "bench cluster=0 exit $?" >> gpurun_out/info.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gate_fused|policy_head" -s 4 -c 4 --csv --log-file gpurun_out/launches_mlp.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
cat gpurun_out/info.log; tail -15 gpurun_out/t_mlp.log; tail -3 gpurun_out/t_model.log; for f in gpurun_out/bench_cl*.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["frac"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
grep -E "gate_fused|policy_head" gpurun_out/launches_mlp.csv | awk -F'","' '{print substr($5,1,40), $9, $15}'
