#!/bin/bash
# round 2, call g (8 GPUs): sharded parity at world 8 (torchrun + single-process context), config 2 scaling (weak + strong),
# and the named shapes of configs 3 and 5 (10^7 g-and-k, 10^8 birth-death particles over 8 GPUs)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 600 $TR --nproc-per-node 8 --master-port 29701 tests/multi_gpu_worker.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -30 ) > gpurun_out/r2g_worker_w8.log
tail -n 18 gpurun_out/r2g_worker_w8.log
( timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k single_process 2>&1 | tail -4 ) > gpurun_out/r2g_multi_ctx.log; cat gpurun_out/r2g_multi_ctx.log
timeout 300 $TR --nproc-per-node 8 --master-port 29718 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2g_weak_n8.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29714 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2g_weak_n4.log 2>&1
for n in 8 2; do
  timeout 300 $TR --nproc-per-node $n --master-port 2972$n bench.py --gpus $n --steps 3 --warmup 2 --particles-total 8000000 > gpurun_out/r2g_strong_n$n.log 2>&1
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_weak_n1.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 2 --particles-total 8000000 --no-cpu-baseline > gpurun_out/r2g_strong_n1.log 2>&1
for f in weak_n1 weak_n4 weak_n8 strong_n1 strong_n2 strong_n8; do tail -n 1 gpurun_out/r2g_$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$f', d['n_gpus'], 'GPUs', d['config']['particles'], 'particles', round(d['value']/1e9,2), 'G evals/s', round(d['ms_per_step'],1), 'ms per run; head us', round(1e3*d['kernels']['head_kernel']['avg_launch_ms'],1), 'sweep us', round(1e3*d['roofline']['avg_launch_ms'],1))
except Exception as e: print('$f FAILED', e)"; done
timeout 600 $TR --nproc-per-node 8 --master-port 29731 bench.py --gpus 8 --config 3 --steps 2 --warmup 1 > gpurun_out/r2g_c3_n8.log 2>&1; tail -n 1 gpurun_out/r2g_c3_n8.log | cut -c1-400
timeout 600 $TR --nproc-per-node 8 --master-port 29732 bench.py --gpus 8 --config 5 --steps 2 --warmup 1 > gpurun_out/r2g_c5_n8.log 2>&1; tail -n 1 gpurun_out/r2g_c5_n8.log | cut -c1-400
timeout 300 $TR --nproc-per-node 8 --master-port 29733 scripts/run_full.py --config 3 --particles-total 10000000 --eps 0.5 --max-iters 60 > gpurun_out/r2g_full_c3.log 2>&1; tail -n 1 gpurun_out/r2g_full_c3.log
timeout 300 $TR --nproc-per-node 8 --master-port 29734 scripts/run_full.py --config 5 --particles-total 100000000 --eps 3.0 --max-iters 40 > gpurun_out/r2g_full_c5.log 2>&1; tail -n 1 gpurun_out/r2g_full_c5.log
timeout 300 $TR --nproc-per-node 8 --master-port 29735 scripts/run_full.py --config 2 --particles-total 100000000 --eps 1.0 > gpurun_out/r2g_full_c2_1e8.log 2>&1; tail -n 1 gpurun_out/r2g_full_c2_1e8.log
