#!/bin/bash
# GPU box: tile mode of the tensor-core GEMM (1 = 128 x 128 per CTA, 2 = CTA pairs 256 x 256) on config 4, forward and training
for md in 1 2; do
  ICNF_TC_MODE=$md python scripts/time_wide.py bf16x3_tc 2>&1 | tail -2
  ICNF_TC_MODE=$md python scripts/time_train4.py bf16x3_tc 2>&1 | tail -2
  ICNF_TC_MODE=$md python scripts/time_wide.py bf16_tc 2>&1 | tail -1
done > gpurun_out/tune_tc.txt 2>&1
if [ "$1" == "ncu" ]; then
for md in 1 2; do
  ICNF_TC_MODE=$md ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/launches_train4_bn$md.csv \
      python scripts/time_train4.py bf16x3_tc > /dev/null 2>&1
done
fi
cat gpurun_out/tune_tc.txt
