#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_v.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_v.log
timeout 900 python tools/sweep_tune.py c5 16 "" "12=1" "0=3" "0=6" 2>&1 | tee gpurun_out/sweep_c5_v.txt
