#!/bin/bash
# GPU box: per-launch device times of the config-4 forward (24 RHS) and training step in both tile modes, and timelines
for md in 1 2; do
  ICNF_TC_MODE=$md ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 130 --csv --log-file gpurun_out/launches_fwd4_m$md.csv \
      python scripts/time_wide.py bf16x3_tc > /dev/null 2>&1
  ICNF_TC_MODE=$md ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_train4_m$md.csv \
      python scripts/time_train4.py bf16x3_tc > /dev/null 2>&1
  ICNF_TC_MODE=$md python scripts/tc_timeline.py 8192 512 512 1 0 > gpurun_out/timeline_act_m$md.txt 2>&1
  ICNF_TC_MODE=$md python scripts/tc_timeline.py 8192 784 512 1 3 > gpurun_out/timeline_plain_m$md.txt 2>&1
done
grep -E "GEMM|acc_ready|epi_done" gpurun_out/timeline_*_m*.txt
