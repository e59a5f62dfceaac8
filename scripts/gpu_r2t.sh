#!/bin/bash
# round 2, call t (1 GPU): FP32 particle state (relaxed mode) -- statistical tests, parity suite of the default path, A/B at 1e6 and 1e7
set -u
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests/ -m gpu -q -x -k "relaxed or fp32 or sweep or run_follows or resample or head" 2>&1 | tail -3
  echo "== config 2, 1e6 particles: parity mode | fp32_state | fp32_state + segments"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-420
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fp32-state 2>&1 | tail -1 | cut -c1-420
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fp32-state --segments 2>&1 | tail -1 | cut -c1-420
  echo "== config 2, 1e7 particles: parity mode | fp32_state"
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --particles 10000000 2>&1 | tail -1 | cut -c1-420
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --particles 10000000 --fp32-state 2>&1 | tail -1 | cut -c1-420
} > gpurun_out/r2t_fp32_state.log 2>&1
cat gpurun_out/r2t_fp32_state.log
