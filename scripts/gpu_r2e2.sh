#!/bin/bash
# round 2, call e2 (2 GPUs, short): single-process multi-GPU context + the torchrun sharded parity
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -30 ) > gpurun_out/r2e_pytest_multi.log
tail -n 14 gpurun_out/r2e_pytest_multi.log
