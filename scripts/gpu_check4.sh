#!/bin/bash
# GPU call 6: validate uniform MMA issue + smem scale/bias + residual prefetch; bench; launch list; nosw probe variants
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
for v in 0 1 2 3; do timeout 30 ./tools/probe_umma_nosw $v >> gpurun_out/probe_nosw.log 2>&1; done
PYT="python -m pytest -m gpu -q --tb=short --timeout 90 --timeout-method=thread"
timeout 600 $PYT tests/test_gpu_kernels.py -k "tc or flat or rowwin or simt" > gpurun_out/k_conv.log 2>&1; echo "k_conv exit $?" > gpurun_out/info.log
timeout 600 $PYT tests/test_gpu_automoe.py > gpurun_out/automoe.log 2>&1; echo "automoe exit $?" >> gpurun_out/info.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/info.log
AMOE_FLAT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_noflat.log 2> gpurun_out/bench_noflat.err; echo "bench(noflat) exit $?" >> gpurun_out/info.log
if grep -q "bench exit 0" gpurun_out/info.log; then
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 230 -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit $?" >> gpurun_out/info.log
fi
cat gpurun_out/info.log; cat gpurun_out/probe_nosw.log
