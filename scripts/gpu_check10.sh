#!/bin/bash
# GPU call 10: L2-chunked stem+layer1, side-stream upsample; A/B over chunk sizes; ncu --set full of stem_pool / flat<64>
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 120 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_automoe.py tests/test_gpu_kernels.py -k "automoe or flat or stem_pool or chunk or bf16 or fp32" > gpurun_out/t_new.log 2>&1; echo "t_new exit $?" > gpurun_out/info.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
for c in 0 16 32 64; do
  AMOE_L2_CHUNK=$c timeout 300 $B > gpurun_out/bench_c$c.log 2> gpurun_out/bench_c$c.err; echo "bench chunk=$c exit $?" >> gpurun_out/info.log
done
AMOE_OVERLAP=0 timeout 300 $B > gpurun_out/bench_noov.log 2> gpurun_out/bench_noov.err; echo "bench no-overlap exit $?" >> gpurun_out/info.log
# source-level profile of the two kernels that bound the first stage (no chunking so the launches are full-size)
export AMOE_L2_CHUNK=0
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel|conv3x3_flat_kernel" -s 10 -c 3 \
    -o gpurun_out/prof3 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/info.log
ncu -i gpurun_out/prof3.ncu-rep --page raw --csv > gpurun_out/prof3_raw.csv 2> gpurun_out/raw.err
for i in 0 1 2; do ncu -i gpurun_out/prof3.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof3_src_$i.csv 2>/dev/null; done
cat gpurun_out/info.log; tail -3 gpurun_out/t_new.log; for f in gpurun_out/bench_c*.log gpurun_out/bench_noov.log; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d["value"]), d["ms_per_step"], d["roofline"]["frac"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
