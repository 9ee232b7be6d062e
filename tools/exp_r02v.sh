#!/bin/bash
set -u
mkdir -p gpurun_out
SWEEP_REPS=3 timeout 900 python tools/sweep_tune.py c5 32 "12=31" "12=34" "12=31,10=24" "12=31,10=16" "12=31,11=12" "12=31,11=8" "12=31,0=6" 2>&1 | tee gpurun_out/sweep_c5_v.txt
