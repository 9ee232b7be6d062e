#!/bin/bash
# round-2 run T: warp-cooperative triangle phase for FLAT scenes in the path kernel (C4)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "state_machine or render_bit_exact or c4_path or path_regeneration or golden or smoke or ragged" > gpurun_out/gputest_t.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputest_t.log
timeout 900 python tools/sweep_tune.py c4 32 "13=2" "" "0=4" "0=10" 2>&1 | tee gpurun_out/sweep_c4_t.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega_path -s 1 -c 1 -o gpurun_out/prof_c4_coop python tools/sweep_tune.py c4 2 "15=23" > gpurun_out/ncu_c4_coop.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_c4_coop.ncu-rep > gpurun_out/prof_c4_coop_summary.txt 2>&1
python tools/ncu_blocks.py gpurun_out/prof_c4_coop.ncu-rep 30 > gpurun_out/prof_c4_coop_blocks.txt 2>&1
head -24 gpurun_out/prof_c4_coop_summary.txt; cat gpurun_out/prof_c4_coop_blocks.txt
