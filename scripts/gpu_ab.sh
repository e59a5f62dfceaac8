#!/bin/bash
# A/B of experiment builds: parity suite on the default build, sweep micro-benchmarks and a short bench for each library
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log
for lib in abcdez.jl_b200/libabcdez_cuda*.so; do
  for args in "gauss_corr10 1000000 0.0" "gauss_corr10 1000000 0.3" "twod 1000000 0.0"; do
    ABCDEZ_LIB=$PWD/$lib timeout 120 python scripts/bench_sweep.py $args 2>&1 | tail -1 | sed "s#.*/libabcdez_cuda##"
  done
  ABCDEZ_LIB=$PWD/$lib timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])"
done > gpurun_out/variants.log
cat gpurun_out/pytest_gpu.log gpurun_out/variants.log
