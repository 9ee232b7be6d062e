"""Builds tests/golden/reference_raycast_b200_opencl.npz from the PPM the UNMODIFIED reference wrote.

Provenance: `gpurun -- bash tools/run_reference_opencl.sh` runs oracle/_ref/adlTest64 (the reference's own
test/RaytraceTest.cpp + ADL + OpenCL backend, built by `make -C oracle ref`) on a B200 through NVIDIA's OpenCL
driver: DeviceTest.RayCast, 512x512, 10000 frames, 357.4 s.  It leaves gpurun_out/reference_opencl/reference.ppm;
this script packs that ASCII P3 file into a compressed uint8 array.

Usage: python tests/golden/make_reference_golden.py [reference.ppm]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "reference_opencl", "reference.ppm")
tok = open(src).read().split()
assert tok[0] == "P3" and tok[3] == "255", tok[:4]
w, h = int(tok[1]), int(tok[2])
rgb = np.array(tok[4:], np.uint8).reshape(h, w, 3)
np.savez_compressed(
    os.path.join(HERE, "reference_raycast_b200_opencl.npz"), rgb=rgb,
    provenance=np.array("UNMODIFIED reference oracle/_ref/adlTest64 --gtest_filter=DeviceTest.RayCast (10000 frames, 512x512) run on an "
                        "NVIDIA B200 through NVIDIA OpenCL 3.0 (driver 580.159) with the -O0 -> -cl-opt-disable interposer; "
                        "tools/run_reference_opencl.sh; 357.4 s"))
print("wrote", rgb.shape, "from", src)
