#!/bin/bash
# round 2, call h (1 GPU): final build -- parity suite, bench line, launch list + ncu --set full of sweep and head (the same
# command as the bench), converged runs of configs 3 and 5 at a size one GPU finishes in about a minute
set -u
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 ) > gpurun_out/r2h_smoke.log; cat gpurun_out/r2h_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.log 2>&1; tail -n 1 gpurun_out/r2h_bench.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.log 2>&1; tail -n 1 gpurun_out/r2h_bench_ref.log | cut -c1-300
for n in 1000000 10000000; do timeout 200 python scripts/bench_sweep.py gauss_corr10 $n 2>&1 | tail -1; done > gpurun_out/r2h_sweep_sizes.log; cat gpurun_out/r2h_sweep_sizes.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2h_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 3 -o gpurun_out/r2h_sweep -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2h_ncu_sweep.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 20 -c 2 -o gpurun_out/r2h_head -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2h_ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gk_smc_sweep -s 2 -c 1 -o gpurun_out/r2h_gk -f python bench.py --config 3 --particles 100000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2h_ncu_gk.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 10 -c 2 -o gpurun_out/r2h_lv -f python bench.py --config 4 --particles 200000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2h_ncu_lv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 4 -c 2 -o gpurun_out/r2h_bd -f python bench.py --config 5 --particles 1000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2h_ncu_bd.log 2>&1
timeout 400 python scripts/run_full.py --config 3 --particles-total 100000 --eps 0.3 > gpurun_out/r2h_full_c3.log 2>&1; tail -n 1 gpurun_out/r2h_full_c3.log
timeout 400 python scripts/run_full.py --config 5 --particles-total 2000000 --eps 1.5 > gpurun_out/r2h_full_c5.log 2>&1; tail -n 1 gpurun_out/r2h_full_c5.log
timeout 400 python scripts/run_full.py --config 4 --particles-total 1000000 > gpurun_out/r2h_full_c4.log 2>&1; tail -n 1 gpurun_out/r2h_full_c4.log
