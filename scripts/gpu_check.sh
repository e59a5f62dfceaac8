#!/bin/bash
# GPU tests + smoke + bench (default and 10^7 particles) + head micro-benchmark
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench.log
( timeout 600 python bench.py --steps 2 --warmup 1 --particles 10000000 --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/bench_1e7.log
timeout 300 python scripts/bench_head.py > gpurun_out/bench_head.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_1e7.log gpurun_out/bench_head.log
