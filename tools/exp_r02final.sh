#!/bin/bash
# round-2 final check: the GPU test suite and smoke() on the committed tree
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
