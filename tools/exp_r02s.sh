#!/bin/bash
# round-2 run S: full GPU test suite + the default bench line with k_path_sm (6 visits per vote, 128 Mi slots per launch)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_s.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_s.log
timeout 900 python bench.py > gpurun_out/bench_c5_n1_s.json 2> gpurun_out/bench_c5_n1_s.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c5_n1_s.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_n1_s.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']))
print('roofline',json.dumps(d['roofline'])[:1800])
for k,v in d['configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), v.get('simt_frac'))
print('clocks',d['clocks'], 'launches', d.get('gpu_launches'))
PY
