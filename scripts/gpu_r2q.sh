#!/bin/bash
# round 2, call q (1 GPU): abcdemc! with the queue-driven sweep for heavy simulators -- parity + A/B against the fused kernel
set -u
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests/ -m gpu -q -x -k "mc or lotka or birth or run_follows or sweep" 2>&1 | tail -3
  for lib in abcdez.jl_b200/libabcdez_cudamcfused.so abcdez.jl_b200/libabcdez_cuda.so; do
    echo "== $lib"
    for c in 4 5; do ABCDEZ_LIB=$PWD/$lib timeout 600 python bench.py --config $c --mc --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330; done
  done
} > gpurun_out/r2q_mc_split.log 2>&1
cat gpurun_out/r2q_mc_split.log
