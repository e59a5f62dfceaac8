#!/bin/bash
# round 2, call w4 (4 GPUs): host-side phase timings of sharded e2e steps (ABCDEZ_TRACE: library phases + bench.py's own clock)
set -u
mkdir -p gpurun_out
ABCDEZ_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2w4_trace_n4.log 2>&1
grep "abcdez\]" gpurun_out/r2w4_trace_n4.log | tail -16; grep "\[bench\]" gpurun_out/r2w4_trace_n4.log; tail -n 1 gpurun_out/r2w4_trace_n4.log | cut -c1-200
