#!/bin/bash
# round-2 run W (final): GPU tests, the default bench line, its ncu launch list, --set full captures of the C5 and C4 path kernels, SASS histograms
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_w.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_w.log
timeout 900 python bench.py > gpurun_out/bench_c5_n1_w.json 2> gpurun_out/bench_c5_n1_w.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5_n1_w.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_n1_w.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']))
print('roofline',json.dumps({k:v for k,v in d['roofline'].items() if k not in ('simt','note','binding')})[:1500])
for k,v in d['configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), v.get('simt_frac'), v.get('e2e') and round(v['e2e']['value']))
print('clocks',d['clocks'], 'launches', d.get('gpu_launches'))
print('cpu',json.dumps(d['cpu_baseline'])[:400])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1_w.json 2> gpurun_out/bench_ref_n1_w.err; tail -c 700 gpurun_out/bench_ref_n1_w.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_w.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_final python tools/sweep_tune.py c5 2 "15=23" > gpurun_out/ncu_c5_final.log 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega_path -s 1 -c 1 -o gpurun_out/prof_c4_final python tools/sweep_tune.py c4 2 "15=23" > gpurun_out/ncu_c4_final.log 2>&1; echo "ncu c4 rc=$?"
for w in c5 c4; do
python tools/ncu_summary.py gpurun_out/prof_${w}_final.ncu-rep > gpurun_out/prof_${w}_final_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_${w}_final.ncu-rep 40 > gpurun_out/prof_${w}_final_blocks.txt 2>&1
ncu -i gpurun_out/prof_${w}_final.ncu-rep --page raw --csv > gpurun_out/prof_${w}_final_raw.csv 2>/dev/null
done
head -30 gpurun_out/prof_c5_final_summary.txt
