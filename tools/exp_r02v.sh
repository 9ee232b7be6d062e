#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "render_bit_exact or golden or c4_path or ragged or state_machine" > gpurun_out/gputest_v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_v.log
SWEEP_REPS=3 timeout 900 python tools/sweep_tune.py c4 32 "13=3" "" "13=3" "" 2>&1 | tee gpurun_out/sweep_c4_v.txt
