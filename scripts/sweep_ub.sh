#!/bin/bash
# Variants of the library differing in the unit-parallel backward's register cap (config-2 shape only).
set -e
cd "$(dirname "$0")/../continuousnormalizingflows.jl_b200/csrc"
mkdir -p ../../sweep/build
cat > ../../sweep/build/inst.cu <<'EOC'
#include "tiny_launch.cuh"
ICNF_REGISTER_TINY(ICNF_ACT_SOFTPLUS, 2, 0, 3, 3, 12, 12, 2)
EOC
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I."
for f in api generic tc; do nvcc $FLAGS -c $f.cu -o ../../sweep/build/$f.o & done
wait
for v in "$@"; do
  ( nvcc $FLAGS -DICNF_UB_MINB=$v -c ../../sweep/build/inst.cu -o ../../sweep/build/inst_ub$v.o && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../sweep/libicnf_b200_ub$v.so ../../sweep/build/api.o ../../sweep/build/generic.o ../../sweep/build/tc.o ../../sweep/build/inst_ub$v.o ) &
done
wait
ls ../../sweep/*.so
