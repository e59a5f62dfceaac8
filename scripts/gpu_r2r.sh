#!/bin/bash
# round 2, call r (8 GPUs): the named shapes of configs 3 and 5 with the final kernels (g-and-k z-space select, queue-driven
# sweeps): 10^7 g-and-k and 10^8 birth-death particles as one sharded population over 8 GPUs
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --config 3 --steps 2 --warmup 1 > gpurun_out/r2r_c3_n8.log 2>&1; tail -n 1 gpurun_out/r2r_c3_n8.log | cut -c1-400
timeout 300 $TR --nproc-per-node 8 --master-port 29732 bench.py --gpus 8 --config 5 --steps 2 --warmup 1 > gpurun_out/r2r_c5_n8.log 2>&1; tail -n 1 gpurun_out/r2r_c5_n8.log | cut -c1-400
timeout 200 $TR --nproc-per-node 8 --master-port 29733 scripts/run_full.py --config 3 --particles-total 10000000 --eps 0.5 --max-iters 160 > gpurun_out/r2r_full_c3.log 2>&1; tail -n 1 gpurun_out/r2r_full_c3.log
timeout 200 $TR --nproc-per-node 8 --master-port 29734 scripts/run_full.py --config 5 --particles-total 100000000 --eps 3.0 --max-iters 100 > gpurun_out/r2r_full_c5.log 2>&1; tail -n 1 gpurun_out/r2r_full_c5.log
