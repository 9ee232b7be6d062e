#!/bin/bash
# Records a round: GPU tests, the default bench line (C5 + C1..C4 sub-records) and the reference arm, the ncu launch list of the
# bench command, --set full captures of the C5 and C4 path kernels with per-block lanes -> gpurun_out/ (copy the summaries into
# profiles/rNN/ afterwards).  usage: bash tools/record_round.sh [tag]     (round 2 ran this as tools/exp_r02w.sh)
set -u
R=${1:-r03}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$R.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_$R.log
timeout 900 python bench.py > gpurun_out/bench_c5_n1_$R.json 2> gpurun_out/bench_c5_n1_$R.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_n1_$R.json 2> gpurun_out/bench_reference_n1_$R.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_$R.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_$R python tools/sweep_tune.py c5 2 "15=23" > gpurun_out/ncu_c5_$R.log 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega_path -s 1 -c 1 -o gpurun_out/prof_c4_$R python tools/sweep_tune.py c4 2 "15=23" > gpurun_out/ncu_c4_$R.log 2>&1; echo "ncu c4 rc=$?"
for w in c5 c4; do
  python tools/ncu_summary.py gpurun_out/prof_${w}_$R.ncu-rep > gpurun_out/prof_${w}_${R}_summary.txt 2>&1
  python tools/ncu_blocks.py gpurun_out/prof_${w}_$R.ncu-rep 40 > gpurun_out/prof_${w}_${R}_blocks.txt 2>&1
  ncu -i gpurun_out/prof_${w}_$R.ncu-rep --page raw --csv > gpurun_out/prof_${w}_${R}_raw.csv 2>/dev/null
done
python - "$R" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_c5_n1_{sys.argv[1]}.json"))
print("C5", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["value"]),
      "roofline", d["roofline"]["bound"], round(d["roofline"]["frac"] or 0, 3), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
for k, v in d["configs"].items():
    print(k, round(v["value"]), "Mrays/s", round(v["ms_per_step"], 4), "ms/step", "simt_frac", round(v.get("simt_frac") or 0, 3))
PY
