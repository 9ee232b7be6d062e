#!/bin/bash
# round-2 run E: GPU tests, then the UNMODIFIED reference through NVIDIA OpenCL with the build / kernel stopwatches
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_e.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/gputest_e.log
bash tools/run_reference_opencl.sh gpurun_out/reference_opencl_r02 2>&1 | tail -40
ls -la gpurun_out/reference_opencl_r02
