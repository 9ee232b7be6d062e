#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "render_multi" > gpurun_out/gputest_v.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/gputest_v.log
