#!/bin/bash
# GPU call 14: detection-expert training step (a12) + re-check of the a11 tests after the conv generalisation
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 1200 $PYT tests/test_gpu_train_det.py tests/test_gpu_train.py > gpurun_out/t_train.log 2>&1; echo "t_train exit $?" > gpurun_out/info.log
cat gpurun_out/info.log; tail -60 gpurun_out/t_train.log
