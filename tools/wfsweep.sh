run() { timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['value']), 'Mrays/s', round(d['ms_per_step'],3), 'ms')"; }
run --workload c4 --integrator wavefront --tune 6=1
run --workload c4 --integrator wavefront --tune 7=4
run --workload c4 --integrator wavefront --tune 7=8
run --workload c4 --integrator wavefront --tune 7=16
run --workload c4 --integrator wavefront --tune 7=24
run --workload c2 --integrator wavefront --tune 6=1 --steps 30
run --workload c2 --integrator wavefront --tune 7=4 --steps 30
run --workload c2 --integrator wavefront --tune 7=8 --steps 30
run --workload c2 --integrator wavefront --tune 7=16 --steps 30
run --workload c5 --integrator wavefront --tune 6=1 --steps 3
run --workload c5 --integrator wavefront --tune 7=8 --steps 3
run --workload c5 --integrator wavefront --tune 7=16 --steps 3
