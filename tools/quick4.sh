for w in c2 c3 c4; do timeout 100 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --ab | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms', {k: round(v['Mrays/s']) for k,v in d['ab'].items()})"; done
for a in "" "--gpu-build"; do timeout 150 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --ab $a | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 $a', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms', {k: round(v['Mrays/s']) for k,v in d['ab'].items()}, round(d['roofline']['nodes_per_ray'],1), round(d['roofline']['tri_tests_per_ray'],2))"; done
