#!/bin/bash
# GPU call 18: source-level profile of the folded stem kernel
mkdir -p gpurun_out
timeout 300 python -m pytest -m gpu -q --tb=short tests/test_gpu_kernels.py -k "stem" > gpurun_out/t_stem.log 2>&1; echo "t_stem exit $?" > gpurun_out/info.log
ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel" -s 2 -c 1 \
    -o gpurun_out/prof6 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/info.log
ncu -i gpurun_out/prof6.ncu-rep --page raw --csv > gpurun_out/prof6_raw.csv 2> gpurun_out/raw.err
ncu -i gpurun_out/prof6.ncu-rep --page source --csv > gpurun_out/prof6_src.csv 2>/dev/null
cat gpurun_out/info.log; tail -2 gpurun_out/t_stem.log
