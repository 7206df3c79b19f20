#!/bin/bash
# ncu evidence for the committed state: complete launch list of one forward + --set full of the tensor-core
# convolutions, the stem and the gate / policy head (one forward)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv|stem|pool|gate|policy|upsample|image_nchw|head1x1|mean_hw" -s 60 -c 90 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
echo "ncu launches exit $?"
ncu --set full --import-source on --clock-control none -k regex:"conv3x3_flat_kernel|conv_tc_kernel|stem_pool|gate_fused|policy_head" -s 27 -c 27 -o gpurun_out/prof_final -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out/prof_final.ncu-rep
