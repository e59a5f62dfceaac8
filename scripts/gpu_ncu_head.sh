#!/bin/bash
# sweep micro-benchmarks of the experiment builds + full ncu captures of the head kernel and the default sweep kernel
set -u
mkdir -p gpurun_out
for lib in abcdez.jl_b200/libabcdez_cuda*.so; do
  for args in "gauss_corr10 1000000 0.0" "gauss_corr10 1000000 0.3" "gauss1d 1000000 0.0" "lotka_volterra 1000000 0.0"; do
    ABCDEZ_LIB=$PWD/$lib timeout 120 python scripts/bench_sweep.py $args 2>&1 | tail -1 | sed "s#.*/libabcdez_cuda##"
  done
done > gpurun_out/variants.log
timeout 300 python scripts/bench_head.py > gpurun_out/bench_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 4 -c 2 -o gpurun_out/prof_head -f python scripts/bench_head.py > gpurun_out/ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 5 -c 2 -o gpurun_out/prof_sweep_micro -f python scripts/bench_sweep.py gauss_corr10 1000000 0.0 > gpurun_out/ncu_micro.log 2>&1
cat gpurun_out/variants.log gpurun_out/bench_head.log
