#!/bin/bash
# round-2 run Z: pooled triangle phase for AO on FLAT scenes (C2)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "render_bit_exact or golden or c2_ao or ragged or single_triangle or sharded" > gpurun_out/gputest_z.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_z.log
SWEEP_REPS=5 timeout 900 python tools/sweep_tune.py c2 1 "13=2" "" "13=2" "" 2>&1 | tee gpurun_out/sweep_c2_z.txt
