#!/bin/bash
# round-2 run J: k_path_sm v2 (one-hot states + REDUX vote, branch-free binary node visit) on C5
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_machine or tessellated or c5_two" > gpurun_out/gputest_j.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_j.log
timeout 900 python tools/sweep_tune.py c5 4 "5=2" "" "10=8" "10=16" "10=20" "11=6" "11=14" "11=18" "0=4" "0=10" "10=16,11=14" "10=8,11=6" 2>&1 | tee gpurun_out/sweep_c5_j.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_sm3 python tools/sweep_tune.py c5 2 "" > gpurun_out/ncu_c5_sm3.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c5_sm3.ncu-rep > gpurun_out/prof_c5_sm3_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c5_sm3.ncu-rep 45 > gpurun_out/prof_c5_sm3_blocks.txt 2>&1
head -24 gpurun_out/prof_c5_sm3_summary.txt; cat gpurun_out/prof_c5_sm3_blocks.txt
