run() { timeout 150 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', round(d['value']), 'Mrays/s', round(d['ms_per_step'],2), 'ms nodes/ray', round(d['roofline']['nodes_per_ray'],1), 'tests/ray', round(d['roofline']['tri_tests_per_ray'],2), 'build', round(d['bvh_build_s'],2))"; }
run --integrator mega --smem-nodes 1024
run --integrator mega --smem-nodes 256
run --integrator mega --smem-nodes 64
run --integrator mega --smem-nodes 8
run --integrator mega --smem-nodes 64 --tune 3=64
run --integrator mega --smem-nodes 64 --max-leaf 2
run --integrator mega --smem-nodes 64 --max-leaf 8
run --integrator wavefront --smem-nodes 64
run --integrator wavefront --smem-nodes 64 --fpb 1
