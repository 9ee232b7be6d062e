#!/bin/bash
# round-2 run K: occupancy sensitivity of k_path_sm on C5 (register caps 9/10 CTAs, lowered occupancy) and BVH leaf-size variants
set -u
mkdir -p gpurun_out
timeout 900 python tools/sweep_tune.py c5 4 "" "12=9" "12=10" "12=1" "9=30" "9=40" 2>&1 | tee gpurun_out/sweep_c5_k.txt
SWEEP_MAX_LEAF=8 timeout 900 python tools/sweep_tune.py c5 4 "" "5=2" 2>&1 | tee -a gpurun_out/sweep_c5_k.txt
SWEEP_SAH_TRAVERSE=2.5 timeout 900 python tools/sweep_tune.py c5 4 "" "5=2" 2>&1 | tee -a gpurun_out/sweep_c5_k.txt
SWEEP_SAH_TRAVERSE=2.5 SWEEP_MAX_LEAF=8 timeout 900 python tools/sweep_tune.py c5 4 "" "5=2" "11=14" 2>&1 | tee -a gpurun_out/sweep_c5_k.txt
SWEEP_SAH_TRAVERSE=5 SWEEP_MAX_LEAF=8 timeout 900 python tools/sweep_tune.py c5 4 "" "5=2" "11=14" 2>&1 | tee -a gpurun_out/sweep_c5_k.txt
