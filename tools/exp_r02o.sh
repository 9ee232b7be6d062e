#!/bin/bash
# round-2 run O: register-cached stack top (k_path_sm v4), 128 Mi sample slots per launch
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_machine or tessellated or c5_two or wavefront" > gpurun_out/gputest_o.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_o.log
timeout 900 python tools/sweep_tune.py c5 16 "" "15=22" "5=2" "10=8" "10=16" "11=6" "11=14" "12=1" "12=9" 2>&1 | tee gpurun_out/sweep_c5_o.txt
SWEEP_INTEGRATOR=wavefront timeout 900 python tools/sweep_tune.py c5 16 "" "13=1" 2>&1 | tee -a gpurun_out/sweep_c5_o.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_c5_sm4 python tools/sweep_tune.py c5 16 "15=23" > gpurun_out/ncu_c5_sm4.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c5_sm4.ncu-rep > gpurun_out/prof_c5_sm4_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c5_sm4.ncu-rep 30 > gpurun_out/prof_c5_sm4_blocks.txt 2>&1
head -24 gpurun_out/prof_c5_sm4_summary.txt; cat gpurun_out/prof_c5_sm4_blocks.txt
