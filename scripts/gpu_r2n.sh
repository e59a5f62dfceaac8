#!/bin/bash
# round 2, call n (1 GPU): g-and-k draws loop, Box-Muller chains in flight per thread (1 / 2 / 4) + ncu of the default
set -u
mkdir -p gpurun_out
{
for lib in abcdez.jl_b200/libabcdez_cudailp1.so abcdez.jl_b200/libabcdez_cuda.so abcdez.jl_b200/libabcdez_cudailp4.so; do
  echo "== $lib"
  ( ABCDEZ_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_runs.py -m gpu -q -x -k "gk_simulate" 2>&1 | tail -1 )
  for m in gk gk_f32; do ABCDEZ_LIB=$PWD/$lib timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1 | sed "s#.*/libabcdez_cuda##"; done
done
} > gpurun_out/r2n_gk_ilp.log 2>&1
cat gpurun_out/r2n_gk_ilp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gk_smc_sweep -s 2 -c 1 -o gpurun_out/r2n_gk -f python bench.py --config 3 --particles 100000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2n_ncu_gk.log 2>&1
tail -2 gpurun_out/r2n_ncu_gk.log | cut -c1-300
