#!/bin/bash
# round 2, call b: the rebuilt head kernel -- parity suite, head / sweep micro-benchmarks, bench, ncu of head + sweep (10^6, 10^7)
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2b_pytest.log
timeout 200 python scripts/bench_head.py 1000000 2>&1 | tail -4 > gpurun_out/r2b_head.log
timeout 200 python scripts/bench_head.py 4000000 2>&1 | tail -4 >> gpurun_out/r2b_head.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2b_bench.log
for n in 1000000 4000000 10000000; do timeout 200 python scripts/bench_sweep.py gauss_corr10 $n 2>&1 | tail -1; done > gpurun_out/r2b_sweep_sizes.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2b_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 20 -c 2 -o gpurun_out/r2b_head -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2b_ncu_head.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 2 -o gpurun_out/r2b_sweep_1e7 -f python bench.py --particles 10000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2b_ncu_1e7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 2 -o gpurun_out/r2b_sweep_1e6 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2b_ncu_1e6.log 2>&1
cat gpurun_out/r2b_pytest.log gpurun_out/r2b_head.log gpurun_out/r2b_bench.log gpurun_out/r2b_sweep_sizes.log
