#!/bin/bash
# round-2 run M: frames per launch (tail amortisation) for k_path_sm and the wavefront on C5
set -u
mkdir -p gpurun_out
for f in 1 2 4 8 16; do SWEEP_FPB=$f timeout 900 python tools/sweep_tune.py c5 16 "" 2>&1 | grep tune | sed "s/^/fpb=$f mega /" | tee -a gpurun_out/sweep_c5_m.txt; done
for f in 2 4 8; do SWEEP_INTEGRATOR=wavefront SWEEP_FPB=$f timeout 900 python tools/sweep_tune.py c5 8 "" "13=1" 2>&1 | grep tune | sed "s/^/fpb=$f wavefront /" | tee -a gpurun_out/sweep_c5_m.txt; done
for f in 1 4 16; do SWEEP_FPB=$f timeout 900 python tools/sweep_tune.py c4 32 "" 2>&1 | grep tune | sed "s/^/fpb=$f mega /" | tee -a gpurun_out/sweep_c5_m.txt; done
