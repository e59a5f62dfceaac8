#!/bin/bash
# round 2, call s (1 GPU): abcde_init! of stepped simulators through the queue -- parity, sanitizers on the queue-driven case, config 5
set -u
mkdir -p gpurun_out
{
  timeout 1500 python -m pytest tests/ -m gpu -q -x -k "init or birth or lotka or run_follows or mc or resume or batch" 2>&1 | tail -3
  for tool in memcheck racecheck; do
    echo "== $tool"; SANITIZE_ONLY="queue-driven" timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py 2>&1 | grep -v "^$" | tail -6
  done
  timeout 300 python scripts/bench_sweep.py birth_death 2000000 2>&1 | tail -1
  timeout 600 python bench.py --config 5 --steps 2 --warmup 2 2>&1 | tail -1
  timeout 300 python scripts/run_full.py --config 5 --particles-total 2000000 --eps 1.5 2>&1 | tail -1
} > gpurun_out/r2s_init_split.log 2>&1
cat gpurun_out/r2s_init_split.log
