#!/bin/bash
# round 2, call j (1 GPU): longest-first queue ordering -- parity of the heavy-simulator paths, birth-death timings
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x -k "birth or lotka or sweep or run_follows or state" 2>&1 | tail -6 ) > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
for n in 200000 1000000 2000000; do timeout 300 python scripts/bench_sweep.py birth_death $n 2>&1 | tail -1; done > gpurun_out/r2j_sweep_bd.log; cat gpurun_out/r2j_sweep_bd.log
timeout 400 python scripts/run_full.py --config 5 --particles-total 2000000 --eps 1.5 > gpurun_out/r2j_full_c5.log 2>&1; tail -n 1 gpurun_out/r2j_full_c5.log | cut -c1-500
timeout 600 python bench.py --config 5 --steps 2 --warmup 1 > gpurun_out/r2j_bench_c5.log 2>&1; tail -n 1 gpurun_out/r2j_bench_c5.log | cut -c1-200
timeout 600 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/r2j_bench_c4.log 2>&1; tail -n 1 gpurun_out/r2j_bench_c4.log | cut -c1-200
timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/r2j_bench_c3.log 2>&1; tail -n 1 gpurun_out/r2j_bench_c3.log | cut -c1-200
