#!/bin/bash
set -u
mkdir -p gpurun_out
SWEEP_REPS=3 timeout 900 python tools/sweep_tune.py c5 32 "" "12=40" "12=29" "12=41" "" "12=40" 2>&1 | tee gpurun_out/sweep_c5_v.txt
