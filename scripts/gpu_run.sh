#!/bin/bash
# One parametrised GPU session (replaces the per-call scripts of round 1).  Usage under gpurun:
#   bash scripts/gpu_run.sh [tests[:<pytest args>]] [smoke] [bench[:<bench args>]] [launches[:<bench args>]] [full:<kernel regex>[:<bench args>]]
# Every stage writes gpurun_out/<stage>.log; ncu stages run only after the same command exited 0 without ncu.
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
PYT="python -m pytest -m gpu -q --tb=short --timeout 600 --timeout-method=thread"
NCU_METRICS="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum"
for stage in "$@"; do
  name="${stage%%:*}"; arg=""; [[ "$stage" == *:* ]] && arg="${stage#*:}"
  case "$name" in
    tests)
      timeout 2400 $PYT ${arg:-tests/} > gpurun_out/tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/info.log; tail -15 gpurun_out/tests.log ;;
    smoke)
      timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/info.log; tail -2 gpurun_out/smoke.log ;;
    bench)
      tag=$(echo "$arg" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)
      timeout 1200 python bench.py $arg > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench [$arg] exit $?" | tee -a gpurun_out/info.log
      tail -c 6000 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err ;;
    launches)
      A="${arg:---steps 1 --warmup 1 --preheat 0 --no-e2e --no-cpu-baseline --no-gpu-reference --no-legs --no-graph}"
      timeout 600 python bench.py $A > gpurun_out/launches_plain.log 2>&1 &&
      timeout 1200 ncu --metrics $NCU_METRICS --clock-control none -k regex:"conv|stem|pool|gate|policy|upsample|image_nchw|stage_u8|head1x1|mean_hw" \
          -s 80 -c 120 --csv --log-file gpurun_out/launches.csv python bench.py $A > gpurun_out/launches_ncu.log 2>&1
      echo "launches exit $?" | tee -a gpurun_out/info.log ;;
    full)
      rx="${arg%%:*}"; A="--steps 1 --warmup 1 --preheat 0 --no-e2e --no-cpu-baseline --no-gpu-reference --no-legs --no-graph"
      [[ "$arg" == *:* ]] && A="${arg#*:}"
      timeout 600 python bench.py $A > gpurun_out/full_plain.log 2>&1 &&
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 40 -c 6 -f -o gpurun_out/full_${rx//[^a-zA-Z0-9]/_} \
          python bench.py $A > gpurun_out/full_ncu.log 2>&1
      echo "full [$rx] exit $?" | tee -a gpurun_out/info.log ;;
    py)
      timeout 1200 python $arg > gpurun_out/py.log 2>&1; echo "py [$arg] exit $?" | tee -a gpurun_out/info.log; tail -30 gpurun_out/py.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
cat gpurun_out/info.log
