#!/bin/bash
# round 2, call m (1 GPU): g-and-k z-space fast path -- parity, sanitizers on the gk case, sweep rate, config 3
set -u
mkdir -p gpurun_out
{
  timeout 900 python -m pytest tests/ -m gpu -q -x -k "gk" 2>&1 | tail -3
  for tool in memcheck racecheck; do
    echo "== $tool"; SANITIZE_ONLY="g-and-k" timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py 2>&1 | grep -v "^$" | tail -8
  done
  for m in gk gk_f32; do timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1; done
  timeout 600 python bench.py --config 3 --steps 3 --warmup 3 2>&1 | tail -1
} > gpurun_out/r2m_gk.log 2>&1
cat gpurun_out/r2m_gk.log
