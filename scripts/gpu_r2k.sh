#!/bin/bash
# round 2, call k (1 GPU): g-and-k CTA size A/B (256 vs 512 threads per particle)
set -u
mkdir -p gpurun_out
for lib in abcdez.jl_b200/libabcdez_cuda.so abcdez.jl_b200/libabcdez_cuda_gk512.so; do
  echo "== $lib"
  ( ABCDEZ_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_runs.py -m gpu -q -x -k "gk" 2>&1 | tail -2 )
  for m in gk gk_f32; do ABCDEZ_LIB=$PWD/$lib timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1 | sed "s#.*/libabcdez_cuda##"; done
done > gpurun_out/r2k_gk_variants.log 2>&1
cat gpurun_out/r2k_gk_variants.log
