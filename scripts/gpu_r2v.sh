#!/bin/bash
# round 2, call v (1 GPU): verification of the final build -- whole GPU suite, smoke, the bench line and its reference arm, launch list +
# ncu --set full of the sweep / head / g-and-k kernels inside their bench commands, bench lines of configs 3-5, converged runs
set -u
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2v_pytest.log; cat gpurun_out/r2v_pytest.log
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 ) > gpurun_out/r2v_smoke.log; cat gpurun_out/r2v_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.log 2>&1; tail -n 1 gpurun_out/r2v_bench.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_ref.log 2>&1; tail -n 1 gpurun_out/r2v_bench_ref.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --fp32-state > gpurun_out/r2v_bench_fp32.log 2>&1; tail -n 1 gpurun_out/r2v_bench_fp32.log | cut -c1-300
for c in 3 4 5; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2v_bench_c$c.log 2>&1; tail -n 1 gpurun_out/r2v_bench_c$c.log | cut -c1-300; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2v_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 3 -o gpurun_out/r2v_sweep -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2v_ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 20 -c 2 -o gpurun_out/r2v_head -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2v_ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gk_smc_sweep -s 2 -c 1 -o gpurun_out/r2v_gk -f python bench.py --config 3 --particles 100000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2v_ncu_gk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 2 -o gpurun_out/r2v_sweep_fp32 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --fp32-state > gpurun_out/r2v_ncu_sweep_fp32.log 2>&1
timeout 400 python scripts/run_full.py --config 3 --particles-total 100000 --eps 0.3 > gpurun_out/r2v_full_c3.log 2>&1; tail -n 1 gpurun_out/r2v_full_c3.log | cut -c1-700
timeout 400 python scripts/run_full.py --config 5 --particles-total 2000000 --eps 1.5 > gpurun_out/r2v_full_c5.log 2>&1; tail -n 1 gpurun_out/r2v_full_c5.log | cut -c1-700
timeout 400 python scripts/run_full.py --config 4 --particles-total 1000000 > gpurun_out/r2v_full_c4.log 2>&1; tail -n 1 gpurun_out/r2v_full_c4.log | cut -c1-700
