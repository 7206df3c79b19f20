#!/bin/bash
# GPU call 16: full GPU suite, default bench (all fields), launch list + dram traffic of every kernel, ncu --set full of stem_pool
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 1500 $PYT tests/ > gpurun_out/all.log 2>&1; echo "all exit $?" > gpurun_out/info.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/info.log
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv|stem|pool|gate|policy|upsample|image_nchw|head1x1" -s 70 -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/info.log
ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel" -s 2 -c 1 \
    -o gpurun_out/prof5 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/info.log
ncu -i gpurun_out/prof5.ncu-rep --page raw --csv > gpurun_out/prof5_raw.csv 2> gpurun_out/raw.err
ncu -i gpurun_out/prof5.ncu-rep --page source --csv > gpurun_out/prof5_src.csv 2>/dev/null
cat gpurun_out/info.log; tail -4 gpurun_out/all.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_default.log
