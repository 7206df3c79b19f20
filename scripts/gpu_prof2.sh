#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel|conv3x3_flat_kernel|gate_fused|policy_head_kernel" -s 27 -c 6 \
    -o gpurun_out/prof2 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" > gpurun_out/info.log
ncu -i gpurun_out/prof2.ncu-rep --page raw --csv > gpurun_out/prof2_raw.csv 2> gpurun_out/raw.err
for i in 0 1 2 3 4 5; do ncu -i gpurun_out/prof2.ncu-rep --page source --csv --launch-skip $i --launch-count 1 > gpurun_out/prof2_src_$i.csv 2>/dev/null; done
ls -la gpurun_out >> gpurun_out/info.log; cat gpurun_out/info.log
