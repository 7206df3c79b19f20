#!/bin/bash
# A/B sweep of environment switches under the sustained (pre-heated, power-capped) bench regime, one box.
#   bash scripts/ab_env.sh "AMOE_FLAT_KW3=1" "AMOE_L2_CHUNK=32" ...      (first run = defaults; defaults again at the end)
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-gpu-reference --no-legs --no-e2e"
run() { env $1 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-28s %9.0f frames/s  %.3f ms  %s MHz %s W' % ('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], round(d['clocks']['power_w_median'])))"; }
run "AMOE_NONE=0" | tee gpurun_out/ab.txt
for e in "$@"; do run "$e" | tee -a gpurun_out/ab.txt; done
run "AMOE_NONE=0" | tee -a gpurun_out/ab.txt
