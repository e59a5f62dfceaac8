#!/bin/bash
# 2+ GPU call: the sharded parity tests, the single-GPU suite as a regression check, 1- and N-GPU bench lines
set -u
NG=${1:-2}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/pytest_multi.log
( timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_head.py > gpurun_out/bench_head.log 2>&1
for m in gk birth_death; do timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1; done > gpurun_out/sweep_micro2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n$NG.log 2>&1
