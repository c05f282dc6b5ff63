#!/bin/bash
# Run on the GPU box (under gpurun): compute-sanitizer over the GPU parity tests of every kernel family.
# memcheck on all of them; racecheck (shared-memory hazards) on the kernels that synchronise through shared memory /
# named barriers.  Summaries land in gpurun_out/sanitizer_*.txt (copied to profiles/ by scripts/collect_profiles.sh).
S=compute-sanitizer
run() {   # name, tool, pytest args
    name=$1; tool=$2; shift 2
    timeout 900 $S --tool $tool --launch-timeout 120 --print-limit 20 python -m pytest "$@" -m gpu -q -x -p no:cacheprovider \
        > gpurun_out/sanitizer_${tool}_${name}.log 2>&1
    { echo "== $tool: pytest $*"; grep -E "passed|failed|error" gpurun_out/sanitizer_${tool}_${name}.log | tail -2;
      grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" gpurun_out/sanitizer_${tool}_${name}.log | sort | uniq -c | head -12; } >> gpurun_out/sanitizer_summary.txt
    rm -f gpurun_out/sanitizer_${tool}_${name}.log.big
}
: > gpurun_out/sanitizer_summary.txt
run tiny memcheck tests/test_gpu_parity.py -k "config2_moons or cond"
run vcabm memcheck tests/test_gpu_vcabm.py
run narrow memcheck tests/test_gpu_narrow.py
run generic memcheck tests/test_gpu_generic.py -k "cond_generic or deep4"
run tc memcheck tests/test_gpu_tc.py -k "selftest or training_gradient_ragged or bf16x3_tc_meets"
run narrow racecheck tests/test_gpu_narrow.py -k "3+0c0-32-32 or generate"
run tiny racecheck tests/test_gpu_parity.py -k "gradient and config2_moons"
run tc racecheck tests/test_gpu_tc.py -k "bf16x3_tc_meets"
run tc_knobs memcheck tests/test_gpu_tc.py -k "knobs"
run planar memcheck tests/test_planar.py
cat gpurun_out/sanitizer_summary.txt
for f in gpurun_out/sanitizer_*.log; do tail -c 20000 $f > $f.tail; rm -f $f; done
