#!/bin/bash
# round 2, call e (1 GPU): head occupancy variants A/B, sanitizers, bench lines of configs 3/4/5
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "head or reweight or quantile" 2>&1 | tail -5 ) > gpurun_out/r2e_pytest_head.log
cat gpurun_out/r2e_pytest_head.log
for lib in abcdez.jl_b200/libabcdez_cuda*.so; do
  echo "== $lib"
  ABCDEZ_LIB=$PWD/$lib timeout 200 python scripts/bench_head.py 1000000 2>&1 | tail -3 | head -2
  ABCDEZ_LIB=$PWD/$lib timeout 200 python scripts/bench_head.py 500000 2>&1 | tail -3 | head -1
  ABCDEZ_LIB=$PWD/$lib timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench', d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['kernels']['head_kernel']['avg_launch_ms'])"
done > gpurun_out/r2e_head_variants.log 2>&1
cat gpurun_out/r2e_head_variants.log
for c in 4 3 5; do timeout 900 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/r2e_bench_c$c.log 2>&1; tail -n 1 gpurun_out/r2e_bench_c$c.log | cut -c1-1500; done
timeout 600 python bench.py --config 3 --model gk_f32 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench_c3_f32.log 2>&1; tail -n 1 gpurun_out/r2e_bench_c3_f32.log | cut -c1-600
timeout 600 python bench.py --config 4 --mc --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench_c4_mc.log 2>&1; tail -n 1 gpurun_out/r2e_bench_c4_mc.log | cut -c1-600
bash scripts/gpu_sanitize.sh
