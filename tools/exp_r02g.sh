#!/bin/bash
# round-2 run G: GPU tests, the default bench line (N=1), its ncu launch list, and --set full captures of the C5 and C4 path kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_g.log
timeout 900 python bench.py > gpurun_out/bench_c5_n1_g.json 2> gpurun_out/bench_c5_n1_g.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5_n1_g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_n1_g.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']))
print('roofline',json.dumps(d['roofline'])[:1500])
for k,v in d['configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), v.get('simt_frac'))
print('clocks',d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_g.log 2>&1; echo "ncu list rc=$?"
for w in c5 c4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega_path -s 2 -c 1 -o gpurun_out/prof_${w}_g python bench.py --workload $w --steps 1 --warmup 1 --spp-per-step 4 --no-sub --no-cpu-baseline --no-e2e > gpurun_out/ncu_${w}_g.log 2>&1; echo "ncu $w rc=$?"
  python tools/ncu_summary.py gpurun_out/prof_${w}_g.ncu-rep > gpurun_out/prof_${w}_g_summary.txt 2>&1
  python tools/ncu_blocks.py gpurun_out/prof_${w}_g.ncu-rep 40 > gpurun_out/prof_${w}_g_blocks.txt 2>&1
done
cat gpurun_out/prof_c5_g_summary.txt
ls -la gpurun_out
