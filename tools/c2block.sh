run() { timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4), 'ms')"; }
run --workload c2
run --workload c2 --tune 3=64
run --workload c2 --tune 3=32
run --workload c3 --tune 3=64
run --workload c3
run --workload c1 --tune 3=64
run --workload c1
