#!/bin/bash
# round-2 run C: GPU tests (full-size parity), the new bench line at N=1 (default = C5, sub-records C1..C4), reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_c.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gputest_c.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_c5_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_n1.json'))
print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']), 'rgb8', d['e2e'] and round(d['e2e']['rgb8']['value']))
print('roofline',{k:d['roofline'].get(k) for k in ('bound','achieved','peak','frac','kernel_ms','kernel_share_of_step')})
print('l2',d['roofline'].get('l2'))
print('simt',d['roofline'].get('simt'))
print('cpu',d['cpu_baseline'])
for k,v in d['configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), v.get('simt_frac'), v.get('e2e'))
print('clocks',d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; tail -c 600 gpurun_out/bench_ref_n1.json
