#!/bin/bash
# round 2, call p (1 GPU): g-and-k with per-precision launch bounds (2 / 3 CTAs per SM)
set -u
mkdir -p gpurun_out
{
  timeout 900 python -m pytest tests/ -m gpu -q -x -k "gk" 2>&1 | tail -3
  for m in gk gk_f32; do timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1; done
  timeout 600 python bench.py --config 3 --steps 3 --warmup 3 2>&1 | tail -1
} > gpurun_out/r2p_gk.log 2>&1
cat gpurun_out/r2p_gk.log
