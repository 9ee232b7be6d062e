run() { timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')"; }
run --workload c4 --integrator mega
run --workload c4 --integrator mega --tune 5=1
run --workload c4 --integrator wavefront
run --workload c5 --integrator mega --steps 3
run --workload c5 --integrator mega --tune 5=1 --steps 3
run --workload c5 --integrator wavefront --steps 3
