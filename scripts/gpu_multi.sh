#!/bin/bash
# 2-GPU call: sharded parity tests + the single-GPU suite
set -u
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/pytest_multi.log
( timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_multi.log; tail -n 30 gpurun_out/pytest_gpu.log
