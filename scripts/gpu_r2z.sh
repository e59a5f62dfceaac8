#!/bin/bash
# round 2, call z (1 GPU): last check of the committed build -- GPU suite, smoke, the default bench line
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/r2z_pytest.log; cat gpurun_out/r2z_pytest.log
( timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2 ) > gpurun_out/r2z_smoke.log; cat gpurun_out/r2z_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench.log 2>&1; tail -n 1 gpurun_out/r2z_bench.log | cut -c1-400
