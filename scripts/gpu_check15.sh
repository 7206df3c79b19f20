#!/bin/bash
# GPU call 15 (2 GPUs): bench under torchrun at N=2 (weak scaling, no collective) + the NCCL training-step check
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?" > gpurun_out/info.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/train_ddp_check.py > gpurun_out/train_ddp.log 2> gpurun_out/train_ddp.err; echo "train ddp exit $?" >> gpurun_out/info.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench reference exit $?" >> gpurun_out/info.log
cat gpurun_out/info.log; tail -2 gpurun_out/bench_n2.log; tail -3 gpurun_out/train_ddp.log; tail -5 gpurun_out/train_ddp.err; tail -1 gpurun_out/bench_ref.log
