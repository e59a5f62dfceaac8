#!/bin/bash
# round 2, call w2 (2 GPUs): the e2e step after the host read of bench.py became a strided sample
set -u
mkdir -p gpurun_out
ABCDEZ_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2w2_trace_n2.log 2>&1
grep "\[bench\]" gpurun_out/r2w2_trace_n2.log; tail -n 1 gpurun_out/r2w2_trace_n2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value ms', d['ms_per_step'], 'e2e', d['e2e'])"
