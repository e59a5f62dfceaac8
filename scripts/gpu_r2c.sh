#!/bin/bash
# round 2, call c: compact head kernel, FP64 g-and-k, device-side abcdemc!: parity suite, micro-benchmarks, ncu of the head
set -u
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r2c_pytest.log
timeout 200 python scripts/bench_head.py 1000000 2>&1 | tail -4 > gpurun_out/r2c_head.log
timeout 200 python scripts/bench_head.py 2000000 2>&1 | tail -4 >> gpurun_out/r2c_head.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2c_bench.log
for m in gk gk_f32 birth_death lotka_volterra; do timeout 300 python scripts/bench_sweep.py $m 200000 2>&1 | tail -1; done > gpurun_out/r2c_sweep_models.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 20 -c 2 -o gpurun_out/r2c_head -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2c_ncu_head.log 2>&1
cat gpurun_out/r2c_pytest.log gpurun_out/r2c_head.log gpurun_out/r2c_bench.log gpurun_out/r2c_sweep_models.log
