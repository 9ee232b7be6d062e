run() { timeout 150 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms')"; }
run --integrator mega
run --integrator mega --tune 2=1
run --integrator mega --tune 2=1 --tune 4=256
run --integrator mega --tune 2=1 --tune 4=1024
run --integrator wavefront --tune 2=1
run --integrator mega --tune 2=1 --gpu-build
