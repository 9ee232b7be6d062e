#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "state_machine or tessellated" > gpurun_out/gputest_v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_v.log
SWEEP_REPS=3 timeout 900 python tools/sweep_tune.py c5 32 "" "11=6" "11=14" 2>&1 | tee gpurun_out/sweep_c5_v.txt
