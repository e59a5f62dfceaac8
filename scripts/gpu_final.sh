#!/bin/bash
# Round-end evidence on one GPU: tests, smoke, bench (both arms), micro-benchmarks, launch list, full ncu captures
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 ) > gpurun_out/bench.log
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 ) > gpurun_out/bench_ref.log
( timeout 300 python scripts/run_breakdown.py 2>&1 | tail -12 ) > gpurun_out/breakdown.log
for m in gauss_corr10 gauss1d twod lotka_volterra; do timeout 120 python scripts/bench_sweep.py $m 1000000 2>&1 | tail -1; done > gpurun_out/sweep_micro.log
for m in gk birth_death; do timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1; done >> gpurun_out/sweep_micro.log
timeout 300 python scripts/bench_head.py > gpurun_out/bench_head.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 3 -o gpurun_out/prof_sweep -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 20 -c 3 -o gpurun_out/prof_head -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_head.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/bench_ref.log gpurun_out/breakdown.log gpurun_out/sweep_micro.log gpurun_out/bench_head.log
ls -la gpurun_out
