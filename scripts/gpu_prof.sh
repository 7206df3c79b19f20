#!/bin/bash
# GPU call: ncu --set full on the conv kernels of one forward (4th forward of a short bench run)
mkdir -p gpurun_out
timeout 60 ./tools/probe_umma_nosw > gpurun_out/probe_nosw.log 2>&1; echo "probe_nosw exit $?" > gpurun_out/info.log
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_flat_kernel|conv_tc_kernel" -s 75 -c 16 \
    -o gpurun_out/prof_conv -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?" >> gpurun_out/info.log
ncu -i gpurun_out/prof_conv.ncu-rep --page raw --csv > gpurun_out/prof_conv_raw.csv 2> gpurun_out/raw.err
ls -la gpurun_out/ >> gpurun_out/info.log
cat gpurun_out/info.log; cat gpurun_out/probe_nosw.log
