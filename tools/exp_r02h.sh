#!/bin/bash
# round-2 run H: the per-lane state-machine path kernel (k_path_sm): parity tests, then quorum-threshold sweeps on C5 and C4
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_machine or tessellated or c5_two or path_regeneration or render_bit_exact" > gpurun_out/gputest_h.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_h.log
timeout 900 python tools/sweep_tune.py c5 4 "5=2" "" "10=4" "10=12" "10=16" "11=4" "11=12" "11=16" "11=2" "0=3" "0=10" "10=12,11=12" "10=16,11=16,0=10" "10=4,11=4,0=4" "4=1" "4=256" 2>&1 | tee gpurun_out/sweep_c5_h.txt
timeout 600 python tools/sweep_tune.py c4 8 "5=2" "5=3" "5=3,10=4" "5=3,10=12" "5=3,10=16" "5=3,11=4" "5=3,11=12" "5=3,11=16" "5=3,10=12,11=12" "5=3,10=16,11=16" "5=3,10=16,11=16,0=12" 2>&1 | tee gpurun_out/sweep_c4_h.txt
