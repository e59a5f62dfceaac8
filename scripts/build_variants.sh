#!/bin/bash
# experiment builds of the sweep kernel: libabcdez_cuda_<tag>.so (see scripts/bench_variants.sh)
set -e
cd "$(dirname "$0")/.."
b() { tag=$1; shift; ABCDEZ_BUILD_TAG=_$tag ABCDEZ_NVCC_EXTRA="-DABCDEZ_SWEEP_CTRL_BATCH=1 -DABCDEZ_SWEEP_PREFETCH=0 $*" python abcdez.jl_b200/build.py > /dev/null; echo "built $tag: $*"; }
b m4 -DABCDEZ_SWEEP_MIN_BLOCKS=4
b xnogather -DABCDEZ_SWEEP_MIN_BLOCKS=4 -DABCDEZ_EXP_NO_GATHER
b xnorng -DABCDEZ_SWEEP_MIN_BLOCKS=4 -DABCDEZ_EXP_NO_RNG
b xneither -DABCDEZ_SWEEP_MIN_BLOCKS=4 -DABCDEZ_EXP_NO_RNG -DABCDEZ_EXP_NO_GATHER
