#!/bin/bash
# One gpurun call: GPU tests on the default build, then the sweep micro-benchmark of every experiment build.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
for lib in abcdez.jl_b200/libabcdez_cuda*.so; do
  for args in "gauss_corr10 1000000 0.0" "gauss_corr10 1000000 0.3" "gauss1d 1000000 0.0"; do
    ABCDEZ_LIB=$PWD/$lib timeout 120 python scripts/bench_sweep.py $args 2>&1 | tail -1 | sed "s#.*/libabcdez_cuda##"
  done
done > gpurun_out/variants.log
( timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench.log
cat gpurun_out/pytest_gpu.log gpurun_out/variants.log gpurun_out/bench.log
