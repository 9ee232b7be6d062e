#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/sweep_tune.py c5 16 "" "14=4" "10=20" 2>&1 | tee gpurun_out/sweep_c5_v.txt
