#!/bin/bash
# round 2, call y (8 GPUs): the relaxed FP32-state mode at scale -- config 2 weak scaling (1e6 per GPU) and 1e8 particles as one population
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29741 bench.py --gpus 8 --steps 5 --warmup 3 --fp32-state --no-cpu-baseline > gpurun_out/r2y_fp32_n8.log 2>&1; tail -n 1 gpurun_out/r2y_fp32_n8.log | cut -c1-330
timeout 200 $TR --nproc-per-node 8 --master-port 29742 bench.py --gpus 8 --steps 1 --warmup 1 --particles 12500000 --fp32-state --no-cpu-baseline > gpurun_out/r2y_fp32_1e8.log 2>&1; tail -n 1 gpurun_out/r2y_fp32_1e8.log | cut -c1-330
