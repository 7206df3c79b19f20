#!/bin/bash
# GPU call 9: 8-warp register-pooled stem, staged/coalesced flat-conv epilogue; tests, bench, launch list
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 90 --timeout-method=thread"
timeout 300 $PYT tests/test_gpu_kernels.py -k "stem_pool or flat" > gpurun_out/k_new.log 2>&1; echo "k_new exit $?" > gpurun_out/info.log
timeout 900 $PYT tests/ > gpurun_out/all.log 2>&1; echo "all exit $?" >> gpurun_out/info.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/info.log
if grep -q "bench exit 0" gpurun_out/info.log; then
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
      -k regex:"conv|stem|pool|gate|policy|upsample|image_nchw|head1x1" -s 160 -c 45 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit $?" >> gpurun_out/info.log
fi
cat gpurun_out/info.log; tail -3 gpurun_out/k_new.log; tail -3 gpurun_out/all.log; cat gpurun_out/bench.log
