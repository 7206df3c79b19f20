#!/bin/bash
# Round-end evidence for the committed state: full GPU suite, smoke, default bench line, launch list of one forward
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 300 --timeout-method=thread"
timeout 1500 $PYT tests/ > gpurun_out/all.log 2>&1; echo "all exit $?" > gpurun_out/info.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/info.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench default exit $?" >> gpurun_out/info.log
timeout 600 python bench.py --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2> gpurun_out/bench_eager.err; echo "bench eager exit $?" >> gpurun_out/info.log
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"conv|stem|pool|gate|policy|upsample|image_nchw|head1x1|mean_hw" -s 60 -c 90 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_run.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/info.log
cat gpurun_out/info.log; tail -3 gpurun_out/all.log; tail -1 gpurun_out/smoke.log; cat gpurun_out/bench_default.log
