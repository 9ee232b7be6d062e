#!/bin/bash
# round-2 run D: the bench line under torchrun (image-sharded strong scaling) at N = $1
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_c5_n$N.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/bench_c5_n{n}.json') if l.startswith('{')][-1])
print('N',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']), 'rgb8', d['e2e'] and round(d['e2e']['rgb8']['value']))
for k,v in d['configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), v.get('simt_frac'), v.get('e2e') and round(v['e2e']['value']))
print('clocks',d['clocks'])
PY
