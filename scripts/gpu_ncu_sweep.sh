#!/bin/bash
# One full ncu capture of the sweep kernel in the micro-benchmark (all particles alive) and one inside a bench run.
set -u
mkdir -p gpurun_out
LIBTAG=${1:-}
export ABCDEZ_LIB=$PWD/abcdez.jl_b200/libabcdez_cuda${LIBTAG}.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 5 -c 2 -o gpurun_out/prof_sweep_micro${LIBTAG} -f python scripts/bench_sweep.py gauss_corr10 1000000 0.0 > gpurun_out/ncu_micro.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 300 -c 2 -o gpurun_out/prof_sweep_run${LIBTAG} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_micro.log gpurun_out/ncu_run.log
ls -la gpurun_out
