#!/bin/bash
set -u
mkdir -p gpurun_out
SWEEP_REPS=3 timeout 600 python tools/sweep_tune.py c4 32 "" "" 2>&1 | tee gpurun_out/sweep_c4_v.txt
