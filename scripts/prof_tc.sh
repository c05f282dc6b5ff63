#!/bin/bash
# GPU box: ncu --set full of tensor-core GEMM launches of the config-4 forward (mode 1: 128 x 128 tiles) and backward
ICNF_TC_MODE=1 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 41 -c 2 -o gpurun_out/prof_tc_fwd_m1 \
    python scripts/time_wide.py bf16x3_tc > /dev/null 2>&1
ICNF_TC_MODE=1 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 300 -c 16 -o gpurun_out/prof_tc_bwd_m1 \
    python scripts/time_train4.py bf16x3_tc > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
