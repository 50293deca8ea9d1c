#!/bin/bash
# Build a library variant with extra -D flags into variants/test_<name>.so (A/B runs: tools/ab.sh name1 name2 ...).
# usage: tools/build_variant.sh <name> [-DRK_... ...]
set -e
name=$1; shift
mkdir -p variants
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -shared -diag-suppress 177 \
    -Xptxas -v "$@" -o variants/test_$name.so rsruckig_b200/csrc/rk_kernels.cu 2> variants/ptxas_$name.log
echo "built variants/test_$name.so"
