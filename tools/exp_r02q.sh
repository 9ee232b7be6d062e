#!/bin/bash
# round-2 run Q: node visits per vote in k_path_sm on C5, with fp32 and quantised nodes
set -u
mkdir -p gpurun_out
export PTB_QNODES=1
timeout 900 python tools/sweep_tune.py c5 16 "14=4" "14=6" "14=20" "14=21" "14=22" "14=6,10=16" "14=6,11=14" "14=6,10=8" "14=6,11=6" "14=6,0=10" 2>&1 | tee gpurun_out/sweep_c5_q.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_sm6 python tools/sweep_tune.py c5 2 "14=6,15=23" > gpurun_out/ncu_c5_sm6.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c5_sm6.ncu-rep > gpurun_out/prof_c5_sm6_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c5_sm6.ncu-rep 30 > gpurun_out/prof_c5_sm6_blocks.txt 2>&1
head -24 gpurun_out/prof_c5_sm6_summary.txt; cat gpurun_out/prof_c5_sm6_blocks.txt
