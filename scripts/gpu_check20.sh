#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 900 $PYT tests/test_gpu_train_det.py tests/test_gpu_automoe.py > gpurun_out/t_new.log 2>&1; echo "t_new exit $?" > gpurun_out/info.log
cat gpurun_out/info.log; tail -30 gpurun_out/t_new.log
