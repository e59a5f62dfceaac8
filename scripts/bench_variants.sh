#!/bin/bash
# time the sweep kernel of every experiment build on a resident population (not a bench.py value)
mkdir -p gpurun_out
for lib in abcdez.jl_b200/libabcdez_cuda_*.so; do
  for args in "gauss_corr10 1000000 0.0" "gauss_corr10 1000000 0.3" "gauss1d 1000000 0.0"; do
    ABCDEZ_LIB=$PWD/$lib timeout 120 python scripts/bench_sweep.py $args 2>&1 | tail -1 | sed "s#.*/libabcdez_cuda_##"
  done
done > gpurun_out/variants.log
