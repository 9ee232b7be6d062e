#!/bin/bash
# round-2 run Y: --set full captures of the C1, C2, C3 kernels (k_mega<PRIMARY|AO|DIRECT>)
set -u
mkdir -p gpurun_out
for w in c1 c2 c3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 2 -c 1 -o gpurun_out/prof_${w}_final python tools/sweep_tune.py $w 1 "" > gpurun_out/ncu_${w}_final.log 2>&1; echo "ncu $w rc=$?"
python tools/ncu_summary.py gpurun_out/prof_${w}_final.ncu-rep > gpurun_out/prof_${w}_final_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_${w}_final.ncu-rep 25 > gpurun_out/prof_${w}_final_blocks.txt 2>&1
head -20 gpurun_out/prof_${w}_final_summary.txt; head -30 gpurun_out/prof_${w}_final_blocks.txt
done
