#!/bin/bash
# 8-GPU call: sharded parity at world 8, then bench lines at N = 8, 4, 2 (N = 1 comes from the single-GPU call)
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 tests/multi_gpu_worker.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -30 ) > gpurun_out/worker_w8.log
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.log 2>&1
done
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.log 2>&1
tail -n 16 gpurun_out/worker_w8.log; for n in 1 2 4 8; do tail -n 1 gpurun_out/scale_n$n.log | cut -c1-400; done
