#!/bin/bash
# GPU call 4: all kernel tests (lock bug fixed), flat conv, full model, bench A/B, ncu launch list.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 90 --timeout-method=thread"
timeout 400 $PYT tests/test_gpu_kernels.py -k "not tc and not flat" > gpurun_out/k_misc.log 2>&1; echo "k_misc exit $?" > gpurun_out/info.log
timeout 600 $PYT tests/test_gpu_kernels.py -k "tc_bf16 or tc_padded or rowwin" > gpurun_out/k_tc.log 2>&1; echo "k_tc exit $?" >> gpurun_out/info.log
timeout 600 $PYT tests/test_gpu_kernels.py -k "flat" > gpurun_out/k_flat.log 2>&1; echo "k_flat exit $?" >> gpurun_out/info.log
timeout 300 $PYT tests/test_gpu_matcher.py > gpurun_out/matcher.log 2>&1; echo "matcher exit $?" >> gpurun_out/info.log
timeout 600 $PYT tests/test_gpu_automoe.py -s > gpurun_out/automoe.log 2>&1; echo "automoe exit $?" >> gpurun_out/info.log
if ! grep -q "automoe exit 0" gpurun_out/info.log; then
  AMOE_FLAT=0 timeout 600 $PYT tests/test_gpu_automoe.py -s > gpurun_out/automoe_noflat.log 2>&1; echo "automoe(AMOE_FLAT=0) exit $?" >> gpurun_out/info.log
  export AMOE_FLAT=0
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/info.log
AMOE_FLAT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_noflat.log 2> gpurun_out/bench_noflat.err; echo "bench(noflat) exit $?" >> gpurun_out/info.log
if grep -q "bench exit 0" gpurun_out/info.log; then
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit $?" >> gpurun_out/info.log
fi
cat gpurun_out/info.log
