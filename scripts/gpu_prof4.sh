#!/bin/bash
mkdir -p gpurun_out
for m in ${MODES:-1}; do
AMOE_MLP_TC=$m ncu --set full --import-source on --clock-control none -k regex:"gate_fused|policy_head" -c 2 -o gpurun_out/prof_mlp_tc$m -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_mlp$m.log 2>&1
echo "ncu tc$m exit $?"
done
ls -la gpurun_out/*.ncu-rep
