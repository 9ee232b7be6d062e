#!/bin/bash
# round-2 run P: prototype of 32-byte quantised binary nodes in k_path_sm on C5
set -u
mkdir -p gpurun_out
export PTB_QNODES=1
timeout 900 python tools/sweep_tune.py c5 16 "" "14=1" "14=2" "14=3" "14=1,11=6" "14=1,11=14" "14=1,10=16" 2>&1 | tee gpurun_out/sweep_c5_p.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_smq python tools/sweep_tune.py c5 2 "14=1,15=23" > gpurun_out/ncu_c5_smq.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c5_smq.ncu-rep > gpurun_out/prof_c5_smq_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c5_smq.ncu-rep 30 > gpurun_out/prof_c5_smq_blocks.txt 2>&1
head -24 gpurun_out/prof_c5_smq_summary.txt; cat gpurun_out/prof_c5_smq_blocks.txt
