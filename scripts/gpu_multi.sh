#!/bin/bash
# 2+ GPU call: the sharded parity tests, then the single-GPU suite as a regression check
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
nvidia-smi topo -m >> gpurun_out/smi_multi.txt 2>&1
( timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/pytest_multi.log
( timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py 2>&1 | tail -8 ) > gpurun_out/pytest_gpu.log
