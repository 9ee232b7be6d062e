#!/bin/bash
# round-2 run F: latency sensitivity of the C5 kernel (occupancy lowered with unused shared memory: tune 9 = KB per CTA),
# staged-prefix size (tune 4), and the C1..C4 sub-records after the single-frame direct write
set -u
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 --spp-per-step 8 --no-sub --no-e2e --no-cpu-baseline "$@" > gpurun_out/f_$tag.json 2> gpurun_out/f_$tag.err; python -c "
import json;d=json.load(open('gpurun_out/f_$tag.json'));print('$tag', round(d['value']), 'Mrays/s', round(d['roofline']['kernel_ms'],3),'ms/launch')"; }
run base
run smem1 --tune 4=1
run smem16 --tune 4=16
run occ6 --tune 9=34
run occ4 --tune 9=52
run occ3 --tune 9=70
run occ2 --tune 9=100
run ld128 --tune 8=1
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/f_full.json 2> gpurun_out/f_full.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_full.json'))
print('c5', round(d['value']))
for k,v in d['configs'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), v.get('simt_frac'))
PY
