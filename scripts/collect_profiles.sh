#!/bin/bash
# Run HERE after `gpurun -- bash scripts/profile_round.sh`: turns what came back in gpurun_out/ into the
# tracked files under profiles/ (ncu summaries, launch lists, bench lines) and regenerates ${R}_numbers.md.
set -e
R=${1:-r02}
cd "$(dirname "$0")/.."
for n in bwd fwd tc tc_x3 generic generic_ws narrow chain tc_train; do
    [ -f gpurun_out/prof_$n.txt ] && cp gpurun_out/prof_$n.txt profiles/${R}_ncu_prof_$n.txt
done
for f in launches_train4.csv time_train4.txt time_config3.txt ab_chain_round.txt; do [ -f gpurun_out/$f ] && cp gpurun_out/$f profiles/${R}_$f; done
cp gpurun_out/bench_full.json profiles/${R}_bench.json
cp gpurun_out/bench_reference.json profiles/${R}_bench_reference.json
cp gpurun_out/launches_bench.csv profiles/${R}_launches_bench.csv
cp gpurun_out/launches_tc.csv profiles/${R}_launches_tc_config4.csv
cp gpurun_out/launches_tc_x3.csv profiles/${R}_launches_tc_x3_config4.csv
cp gpurun_out/launches_c3.csv profiles/${R}_launches_config3.csv
cp gpurun_out/time_wide.txt profiles/${R}_time_wide.txt
for n in 2 4 8; do [ -f gpurun_out/bench_n$n.json ] && tail -1 gpurun_out/bench_n$n.json > profiles/${R}_bench_${n}gpu.json; done
R=$R python - <<'PY'
import json, re, os
R = os.environ["R"]
t = open(f"profiles/{R}_ncu_prof_bwd.txt").read()
rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+) (\w+)", t).group(1))
unit = re.search(r"dram__bytes_read.sum\s+([\d.]+) (\w+)", t).group(2)
wr = re.search(r"dram__bytes_write.sum\s+([\d.]+) (\w+)", t)
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = rd * mul[unit] + float(wr.group(1)) * mul[wr.group(2)]
json.dump({"tiny::backward_sp_kernel": {"dram_bytes_per_launch": int(tot), "batch": 65536,
           "source": "ncu --set full (cache control: flush), dram__bytes_read.sum + dram__bytes_write.sum of one launch, "
                     f"profiles/{R}_ncu_prof_bwd.txt; the launch reads the stage-input checkpoints (48 B per sample and "
                     "accepted step, 4-5 steps) and nothing else of size"}}, open(f"profiles/{R}_traffic.json", "w"), indent=1)
PY
python scripts/make_profiles_numbers.py $R > /dev/null
echo "profiles/ refreshed"
