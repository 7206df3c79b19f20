#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"kw3" -s 1 -c 2 -o gpurun_out/prof_kw3 -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_kw3.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_kw3.ncu-rep
