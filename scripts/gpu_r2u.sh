#!/bin/bash
# round 2, call u (2 GPUs): sharded parity on the final build (queue-driven init / sweeps / abcdemc! across ranks, run-state snapshots,
# single-process multi-GPU context), one 2-GPU bench line, one 2-GPU FP32-state run
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^W\|^\*\*\*" | tail -45 ) > gpurun_out/r2u_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2u_bench_n2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --fp32-state --no-cpu-baseline > gpurun_out/r2u_bench_n2_fp32.log 2>&1
tail -n 32 gpurun_out/r2u_pytest_multi.log; for f in r2u_bench_n2 r2u_bench_n2_fp32; do tail -n 1 gpurun_out/$f.log | cut -c1-600; done
