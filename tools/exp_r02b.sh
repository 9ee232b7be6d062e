#!/bin/bash
# round-2 experiment B: FLAT form on the Cornell box (force-width 4 = the round-1 4-wide tree) on C1..C4
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_b.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gputest_b.log
for w in c4 c2 c3 c1; do for fw in 0 4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --force-width $fw > gpurun_out/${w}_fw$fw.json 2> gpurun_out/${w}_fw$fw.err
  python -c "
import json;d=json.load(open('gpurun_out/${w}_fw$fw.json'));print('$w force_width=$fw', round(d['value']), 'Mrays/s', round(d['ms_per_step'],4),'ms', d['roofline']['nodes_per_ray'], d['roofline']['tri_tests_per_ray'])"; done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega_path -s 13 -c 1 -o gpurun_out/prof_c4_flat python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4_flat.log 2>&1
ncu -i gpurun_out/prof_c4_flat.ncu-rep --page raw --csv > gpurun_out/prof_c4_flat_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_c4_flat.ncu-rep > gpurun_out/prof_c4_flat_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c4_flat.ncu-rep 40 > gpurun_out/prof_c4_flat_blocks.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 13 -c 1 -o gpurun_out/prof_c2_flat python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c2_flat.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_c2_flat.ncu-rep > gpurun_out/prof_c2_flat_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c2_flat.ncu-rep 40 > gpurun_out/prof_c2_flat_blocks.txt 2>&1
cat gpurun_out/prof_c4_flat_summary.txt | head -40
