#!/bin/bash
# round-2: the one-process multi-GPU C/C++ path on N devices: the tests that skip on one GPU, then the reference's RayCast flow (host/ptb_raycast) on 1 and N GPUs
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "${SKIP_TESTS:-0}" = 0 ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "render_multi or deals_frame_ahead or gather or sharded or launch1d or unmodified_reference or cpp_raycast" > gpurun_out/gputest_multi_$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gputest_multi_$N.log; fi
for g in 1 $N; do
  t0=$(date +%s%N); oclpathtracer_b200/host/ptb_raycast data/cornellbox.bin gpurun_out/raycast_g$g.ppm 512 10000 $g 2>&1 | tail -2; t1=$(date +%s%N)
  echo "ptb_raycast 512^2 x 10000 frames gpus=$g wall $(( (t1 - t0) / 1000000 )) ms" | tee -a gpurun_out/raycast_multi_$N.txt
done
cmp gpurun_out/raycast_g1.ppm gpurun_out/raycast_g$N.ppm && echo "512^2 PPM identical on 1 and $N GPUs" | tee -a gpurun_out/raycast_multi_$N.txt
for g in 1 $N; do
  t0=$(date +%s%N); oclpathtracer_b200/host/ptb_raycast data/cornellbox.bin gpurun_out/raycast_big_g$g.ppm 2048 2000 $g 2>&1 | tail -2; t1=$(date +%s%N)
  echo "ptb_raycast 2048^2 x 2000 frames gpus=$g wall $(( (t1 - t0) / 1000000 )) ms" | tee -a gpurun_out/raycast_multi_$N.txt
done
cmp gpurun_out/raycast_big_g1.ppm gpurun_out/raycast_big_g$N.ppm && echo "2048^2 PPM identical on 1 and $N GPUs" | tee -a gpurun_out/raycast_multi_$N.txt
rm -f gpurun_out/raycast_big_g*.ppm gpurun_out/raycast_g*.ppm
