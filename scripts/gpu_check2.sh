#!/bin/bash
# GPU call 3: init-order diagnostic, then everything with our library initialised first.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 700 python scripts/diag_order.py > gpurun_out/order.log 2>&1; echo "order exit $?" > gpurun_out/info.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short -k "simt" --timeout 60 --timeout-method=thread > gpurun_out/k_simt.log 2>&1; echo "k_simt exit $?" >> gpurun_out/info.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=line -k "tc_bf16" --timeout 60 --timeout-method=thread > gpurun_out/k_tc.log 2>&1; echo "k_tc exit $?" >> gpurun_out/info.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -k "rowwin" --timeout 60 --timeout-method=thread > gpurun_out/k_rowwin.log 2>&1; echo "k_rowwin exit $?" >> gpurun_out/info.log
timeout 600 python -m pytest tests/test_gpu_automoe.py -m gpu -q --tb=short --timeout 200 --timeout-method=thread -s > gpurun_out/automoe.log 2>&1; echo "automoe exit $?" >> gpurun_out/info.log
if ! grep -q "automoe exit 0" gpurun_out/info.log; then
  AMOE_STEM=simt timeout 600 python -m pytest tests/test_gpu_automoe.py -m gpu -q --tb=short --timeout 200 --timeout-method=thread -s > gpurun_out/automoe_simtstem.log 2>&1; echo "automoe(simt stem) exit $?" >> gpurun_out/info.log
  export AMOE_STEM=simt
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/info.log
cat gpurun_out/info.log; tail -3 gpurun_out/bench.log
if grep -q "bench exit 0" gpurun_out/info.log; then
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit $?" >> gpurun_out/info.log
fi
cat gpurun_out/info.log
