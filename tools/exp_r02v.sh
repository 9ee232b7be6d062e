#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/sweep_tune.py c3 64 "" 2>&1 | tee gpurun_out/sweep_c3_v.txt
SWEEP_REPS=5 timeout 900 python tools/sweep_tune.py c3 64 "" 2>&1 | tee -a gpurun_out/sweep_c3_v.txt
