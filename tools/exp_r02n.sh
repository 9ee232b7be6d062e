#!/bin/bash
# round-2 run N: 4-wide nodes from global memory in k_path_sm (prototype) on C5
set -u
mkdir -p gpurun_out
export PTB_WIDE_LARGE=1
SWEEP_FPB=16 timeout 900 python tools/sweep_tune.py c5 16 "" "14=3" "14=1" "14=3,11=6" "14=3,11=14" "14=3,10=8" "14=3,10=16" "5=2" 2>&1 | tee gpurun_out/sweep_c5_n.txt
SWEEP_FPB=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_sm4 python tools/sweep_tune.py c5 16 "14=3" > gpurun_out/ncu_c5_sm4.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c5_sm4.ncu-rep > gpurun_out/prof_c5_sm4_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c5_sm4.ncu-rep 30 > gpurun_out/prof_c5_sm4_blocks.txt 2>&1
head -24 gpurun_out/prof_c5_sm4_summary.txt; cat gpurun_out/prof_c5_sm4_blocks.txt
