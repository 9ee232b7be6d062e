#!/bin/bash
# Records the round's bench lines and ncu evidence into gpurun_out/ (copy the summaries into profiles/ afterwards).
set -u
R=${1:-r01}
timeout 300 python bench.py > gpurun_out/bench_c2_$R.json 2> gpurun_out/bench_c2_$R.err
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_$R.json 2>> gpurun_out/bench_c2_$R.err
for w in c1 c3 c4; do timeout 200 python bench.py --workload $w --steps 20 --warmup 3 --ab > gpurun_out/bench_${w}_$R.json 2> gpurun_out/bench_${w}_$R.err; done
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 --ab > gpurun_out/bench_c5_$R.json 2> gpurun_out/bench_c5_$R.err
timeout 100 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$R.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$R.log 2>&1
timeout 100 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain2_$R.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 4 -c 1 -o gpurun_out/prof_mega_ao_final_$R python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$R.log 2>&1
for f in gpurun_out/bench_*_$R.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "ab", {k: round(v["Mrays/s"]) for k,v in (d.get("ab") or {}).items()}, "clk", d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("samples"), d.get("clocks",{}).get("reasons"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
