#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"policy_head_kernel" -s 2 -c 1 \
    -o gpurun_out/prof7 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" > gpurun_out/info.log
ncu -i gpurun_out/prof7.ncu-rep --page raw --csv > gpurun_out/prof7_raw.csv 2> gpurun_out/raw.err
ncu -i gpurun_out/prof7.ncu-rep --page source --csv > gpurun_out/prof7_src.csv 2>/dev/null
cat gpurun_out/info.log
