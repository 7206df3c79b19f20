#!/bin/bash
# One gpurun call: kernel parity tests, full-model parity, matcher, bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/info.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "not tc" --timeout 300 > gpurun_out/t1.log 2>&1
echo "t1 exit $?" >> gpurun_out/info.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "tc" --timeout 120 > gpurun_out/t2.log 2>&1
T2=$?
echo "t2 exit $T2" >> gpurun_out/info.log
if [ $T2 -eq 0 ]; then
  timeout 1200 python -m pytest tests/test_gpu_matcher.py tests/test_gpu_automoe.py -m gpu -q --timeout 600 -s > gpurun_out/t3.log 2>&1
  echo "t3 exit $?" >> gpurun_out/info.log
  timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/info.log
else
  timeout 900 python -m pytest tests/test_gpu_matcher.py tests/test_gpu_automoe.py -m gpu -q --timeout 600 -k "matcher or fp32" > gpurun_out/t3.log 2>&1
  echo "t3(fp32 only) exit $?" >> gpurun_out/info.log
fi
tail -5 gpurun_out/t1.log gpurun_out/t2.log gpurun_out/t3.log
cat gpurun_out/info.log
