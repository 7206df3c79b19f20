#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 60 ./tools/probe_umma > gpurun_out/probe.log 2>&1; echo "probe exit $?" > gpurun_out/info.log
timeout 1300 python scripts/diag_conv.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/info.log
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --tb=short -k "not conv" --timeout 100 --timeout-method=thread > gpurun_out/t1.log 2>&1; echo "t1 exit $?" >> gpurun_out/info.log
cat gpurun_out/probe.log | head -50; cat gpurun_out/diag.log; tail -30 gpurun_out/t1.log; cat gpurun_out/info.log
