#!/bin/bash
# round-2 run U: quantised 32-byte binary nodes in every large-scene kernel: full GPU tests, C5 sweep, ncu
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_u.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gputest_u.log
timeout 900 python tools/sweep_tune.py c5 16 "" "14=4" "14=2" "12=9" "12=1" "10=16" "11=14" "11=7" "5=2" 2>&1 | tee gpurun_out/sweep_c5_u.txt
SWEEP_INTEGRATOR=wavefront timeout 900 python tools/sweep_tune.py c5 16 "" "13=1" 2>&1 | tee -a gpurun_out/sweep_c5_u.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_q python tools/sweep_tune.py c5 2 "15=23" > gpurun_out/ncu_c5_q.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c5_q.ncu-rep > gpurun_out/prof_c5_q_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c5_q.ncu-rep 30 > gpurun_out/prof_c5_q_blocks.txt 2>&1
head -24 gpurun_out/prof_c5_q_summary.txt; head -20 gpurun_out/prof_c5_q_blocks.txt
