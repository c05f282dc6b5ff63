#!/bin/bash
# Builds variants of the library that differ only in the __launch_bounds__ min-blocks of
# the tiny kernels (config-2 shape only, to keep compile time short) into sweep/.
# Run the bench against each with ICNF_B200_LIB=sweep/libicnf_b200_fXbY.so.
set -e
cd "$(dirname "$0")/../continuousnormalizingflows.jl_b200/csrc"
mkdir -p ../../sweep/build
cat > ../../sweep/build/inst.cu <<'EOC'
#include "tiny_launch.cuh"
ICNF_REGISTER_TINY(ICNF_ACT_SOFTPLUS, 2, 0, 3, 3, 12, 12, 2)
EOC
FLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I."
nvcc $FLAGS -c api.cu -o ../../sweep/build/api.o &
nvcc $FLAGS -c generic.cu -o ../../sweep/build/generic.o &
wait
for v in "$@"; do
  f=${v%,*}; b=${v#*,}
  ( nvcc $FLAGS -DICNF_TINY_MINB_FWD=$f -DICNF_TINY_MINB_BWD=$b -c ../../sweep/build/inst.cu -o ../../sweep/build/inst_f${f}b${b}.o && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../sweep/libicnf_b200_f${f}b${b}.so ../../sweep/build/api.o ../../sweep/build/generic.o ../../sweep/build/inst_f${f}b${b}.o ) &
done
wait
ls -la ../../sweep/*.so
