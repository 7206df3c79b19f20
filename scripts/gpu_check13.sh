#!/bin/bash
# GPU call 13: training-step kernels (a11) parity
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 180 --timeout-method=thread"
timeout 1200 $PYT tests/test_gpu_train.py > gpurun_out/t_train.log 2>&1; echo "t_train exit $?" > gpurun_out/info.log
cat gpurun_out/info.log; tail -60 gpurun_out/t_train.log
