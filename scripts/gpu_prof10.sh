#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"head1x1" -s 3 -c 2 -o gpurun_out/prof_h1 -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_h1.log 2>&1
echo "ncu exit $?"
