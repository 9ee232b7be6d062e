#!/bin/bash
# round-2 run R: latency sensitivity of k_path_sm on C5 (resident CTAs per SM)
set -u
mkdir -p gpurun_out
export PTB_QNODES=1
timeout 900 python tools/sweep_tune.py c5 16 "14=6" "14=6,9=-7" "14=6,9=-6" "14=6,9=-5" "14=6,9=-4" "14=21" "14=21,9=-7" "14=21,9=-6" 2>&1 | tee gpurun_out/sweep_c5_r.txt
