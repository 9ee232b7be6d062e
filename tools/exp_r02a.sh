#!/bin/bash
# round-2 experiment A: 256-bit node loads on the 2M-triangle scene (tune 8=1 restores four 128-bit loads)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_a.log
for t in 0 1 0 1; do timeout 300 python bench.py --workload c5 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --tune 8=$t > gpurun_out/c5_ld_$t.json 2> gpurun_out/c5_ld_$t.err; python -c "
import json;d=json.load(open('gpurun_out/c5_ld_$t.json'));print('tune8=$t', round(d['value']), 'Mrays/s', d['ms_per_step'])"; done
for t in 0 1; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega_path -s 2 -c 1 -o gpurun_out/prof_c5_ld_$t python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --tune 8=$t > gpurun_out/ncu_c5_ld_$t.log 2>&1
  ncu -i gpurun_out/prof_c5_ld_$t.ncu-rep --page raw --csv > gpurun_out/prof_c5_ld_${t}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_c5_ld_$t.ncu-rep --page details > gpurun_out/prof_c5_ld_${t}_details.txt 2>/dev/null
done
ls -la gpurun_out | head -40
