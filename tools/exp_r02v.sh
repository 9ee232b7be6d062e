#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "golden_vectors or quantised" > gpurun_out/gputest_v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_v.log
