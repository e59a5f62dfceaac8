#!/bin/bash
# tests + bench + launch list (one GPU)
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench.log
timeout 300 python scripts/bench_head.py > gpurun_out/bench_head.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/bench.log gpurun_out/bench_head.log
