#!/bin/bash
# Run on the GPU box (under gpurun): the bench, the reference arm, the ncu launch list of the
# bench command and one full ncu capture per dominant kernel.  Outputs land in gpurun_out/.
set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_full.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:backward_sp -s 3 -c 1 -o gpurun_out/prof_bwd \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_bwd.ncu-rep > gpurun_out/prof_bwd.txt 2>&1; rm -f gpurun_out/prof_bwd.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:solve_adaptive -s 3 -c 1 -o gpurun_out/prof_fwd \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_fwd.ncu-rep > gpurun_out/prof_fwd.txt 2>&1; rm -f gpurun_out/prof_fwd.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 40 -c 1 -o gpurun_out/prof_tc \
    python scripts/time_wide.py bf16_tc > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_tc.ncu-rep > gpurun_out/prof_tc.txt 2>&1; rm -f gpurun_out/prof_tc.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:gemm128_kernel -s 40 -c 1 -o gpurun_out/prof_generic \
    python scripts/time_wide.py fp32 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_generic.ncu-rep > gpurun_out/prof_generic.txt 2>&1; rm -f gpurun_out/prof_generic.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 40 -c 1 -o gpurun_out/prof_tc_x3 \
    python scripts/time_wide.py bf16x3_tc > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_tc_x3.ncu-rep > gpurun_out/prof_tc_x3.txt 2>&1; rm -f gpurun_out/prof_tc_x3.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 150 --csv --log-file gpurun_out/launches_tc.csv \
    python scripts/time_wide.py bf16_tc > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/launches_tc_x3.csv \
    python scripts/time_wide.py bf16x3_tc > /dev/null 2>&1
python scripts/time_wide.py fp32 > gpurun_out/time_wide.txt 2>&1
python scripts/time_wide.py bf16_tc >> gpurun_out/time_wide.txt 2>&1
python scripts/time_wide.py bf16x3_tc >> gpurun_out/time_wide.txt 2>&1
ICNF_NARROW=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_c3.csv \
    python scripts/time_config3.py fp32 > /dev/null 2>&1
ICNF_NARROW=0 ncu --set full --clock-control none --import-source on -k regex:gemm_ws -s 41 -c 1 -o gpurun_out/prof_generic_ws \
    python scripts/time_config3.py fp32 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_generic_ws.ncu-rep > gpurun_out/prof_generic_ws.txt 2>&1; rm -f gpurun_out/prof_generic_ws.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 330 --csv --log-file gpurun_out/launches_train4.csv \
    python scripts/time_train4.py bf16x3_tc > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 60 -c 2 -o gpurun_out/prof_tc_train \
    python scripts/time_train4.py bf16x3_tc > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_tc_train.ncu-rep > gpurun_out/prof_tc_train.txt 2>&1; rm -f gpurun_out/prof_tc_train.ncu-rep
python scripts/time_train4.py bf16x3_tc > gpurun_out/time_train4.txt 2>&1
ls -la gpurun_out
ncu --set full --clock-control none --import-source on -k regex:solve_kernel -s 1 -c 1 -o gpurun_out/prof_narrow \
    python scripts/time_config3.py fp32 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_narrow.ncu-rep > gpurun_out/prof_narrow.txt 2>&1; rm -f gpurun_out/prof_narrow.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 30 -c 1 -o gpurun_out/prof_chain \
    python scripts/time_wide.py bf16x3_tc > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/prof_chain.ncu-rep > gpurun_out/prof_chain.txt 2>&1; rm -f gpurun_out/prof_chain.ncu-rep
python scripts/time_config3.py fp32 > gpurun_out/time_config3.txt 2>&1
python scripts/ab_chain.py c4 8 adaptive > gpurun_out/ab_chain_round.txt 2>&1
