#!/bin/bash
# round 2, call d (2 GPUs): sharded parity with the all-CTA exchanges of the new head, 2-GPU bench (weak + strong), head micro-benchmark
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/r2d_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2d_bench_n2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --particles-total 1000000 > gpurun_out/r2d_bench_n2_strong.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2d_bench_n1.log 2>&1
timeout 200 python scripts/bench_head.py 1000000 2>&1 | tail -4 > gpurun_out/r2d_head.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2d_ref_n2.log 2>&1
tail -n 25 gpurun_out/r2d_pytest_multi.log; for f in r2d_bench_n1 r2d_bench_n2 r2d_bench_n2_strong r2d_ref_n2; do tail -n 1 gpurun_out/$f.log | cut -c1-3000; done; cat gpurun_out/r2d_head.log
