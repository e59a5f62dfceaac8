#!/bin/bash
# round 2, call a: sanity of the round-1 build on this pool + sweep kernel at 10^6 vs 10^7 particles under ncu --set full
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2a_pytest.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2a_bench.log
for n in 1000000 4000000 10000000; do timeout 200 python scripts/bench_sweep.py gauss_corr10 $n 2>&1 | tail -1; done > gpurun_out/r2a_sweep_sizes.log
timeout 200 python scripts/bench_head.py 1000000 2>&1 | tail -4 > gpurun_out/r2a_head.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 2 -o gpurun_out/r2a_sweep_1e7 -f python bench.py --particles 10000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2a_ncu_1e7.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_sweep_kernel -s 20 -c 2 -o gpurun_out/r2a_sweep_1e6 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2a_ncu_1e6.log 2>&1
cat gpurun_out/r2a_pytest.log gpurun_out/r2a_bench.log gpurun_out/r2a_sweep_sizes.log gpurun_out/r2a_head.log
