#!/bin/bash
set -u
mkdir -p gpurun_out
SWEEP_REPS=4 timeout 900 python tools/sweep_tune.py c5 32 "" "12=29" "12=31" "" "12=29" "12=31" 2>&1 | tee gpurun_out/sweep_c5_v.txt
