#!/bin/bash
# ncu launch list (durations) of one eager forward, all kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv|stem|pool|gate|policy|upsample|image_nchw|head1x1|mean" -s 60 -c 90 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
echo "ncu launches exit $?"
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/fwd_breakdown.csv gpurun_out/conv_traffic.json; cat gpurun_out/fwd_breakdown.csv | cut -c1-120
