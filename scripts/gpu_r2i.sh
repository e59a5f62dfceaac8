#!/bin/bash
# round 2, call i (1 GPU): queue-driven sweeps of the heavy simulators (lane refill for the birth-death SSA, dense warps for
# Lotka-Volterra): parity suite, sweep micro-benchmarks, bench lines of configs 4 and 5, ncu of the simulate kernel
set -u
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 ) > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
for m in birth_death lotka_volterra; do for n in 200000 2000000; do timeout 300 python scripts/bench_sweep.py $m $n 2>&1 | tail -1; done; done > gpurun_out/r2i_sweep_models.log; cat gpurun_out/r2i_sweep_models.log
timeout 600 python bench.py --config 5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_bench_c5.log 2>&1; tail -n 1 gpurun_out/r2i_bench_c5.log | cut -c1-200
timeout 600 python bench.py --config 4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_bench_c4.log 2>&1; tail -n 1 gpurun_out/r2i_bench_c4.log | cut -c1-200
timeout 400 python scripts/run_full.py --config 5 --particles-total 2000000 --eps 1.5 > gpurun_out/r2i_full_c5.log 2>&1; tail -n 1 gpurun_out/r2i_full_c5.log | cut -c1-700
timeout 400 python scripts/run_full.py --config 4 --particles-total 1000000 > gpurun_out/r2i_full_c4.log 2>&1; tail -n 1 gpurun_out/r2i_full_c4.log | cut -c1-700
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simulate_queue_kernel -s 4 -c 1 -o gpurun_out/r2i_bd_queue -f python bench.py --config 5 --particles 1000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2i_ncu_bd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simulate_queue_kernel -s 10 -c 1 -o gpurun_out/r2i_lv_queue -f python bench.py --config 4 --particles 200000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2i_ncu_lv.log 2>&1
