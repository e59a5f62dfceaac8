#!/bin/bash
# round 2, call w (2 GPUs): where does the e2e overhead of sharded runs go?  host-side phase timings of abcdez_smc_run (ABCDEZ_TRACE)
set -u
mkdir -p gpurun_out
ABCDEZ_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2w_trace_n2.log 2>&1
grep "abcdez\]" gpurun_out/r2w_trace_n2.log | tail -24; tail -n 1 gpurun_out/r2w_trace_n2.log | cut -c1-200
ABCDEZ_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2w_trace_n1.log 2>&1
grep "abcdez\]" gpurun_out/r2w_trace_n1.log | tail -8
