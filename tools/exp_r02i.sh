#!/bin/bash
# round-2 run I: k_path_sm vs k_mega_path_regen (tune[5] = 3 / 2) on C5 and C4, ncu of k_path_sm on C5
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_machine" > gpurun_out/gputest_i.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_i.log
timeout 900 python tools/sweep_tune.py c5 4 "5=2" "5=3" "5=1" 2>&1 | tee gpurun_out/sweep_c5_i.txt
timeout 600 python tools/sweep_tune.py c4 8 "5=2" "5=3" "5=3,10=4" "5=3,10=16" "5=3,11=4" "5=3,11=16"  "5=3,10=16,11=16" 2>&1 | tee gpurun_out/sweep_c4_i.txt
for w in c5 c4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_path_sm -s 1 -c 1 -o gpurun_out/prof_${w}_sm python tools/sweep_tune.py $w 2 "5=3" > gpurun_out/ncu_${w}_sm.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_${w}_sm.ncu-rep > gpurun_out/prof_${w}_sm_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_${w}_sm.ncu-rep 45 > gpurun_out/prof_${w}_sm_blocks.txt 2>&1
head -24 gpurun_out/prof_${w}_sm_summary.txt; cat gpurun_out/prof_${w}_sm_blocks.txt
done
