#!/bin/bash
# Runs the UNMODIFIED reference (oracle/_ref/adlTest64 = its DeviceTest.RayCast: 10000 frames, 512x512) on this box's
# GPU through NVIDIA's OpenCL driver, then the same flow through libptb200.so, and compares the two PPMs.
# The box ships libnvidia-opencl.so.1 but no ICD registry entry; the entry is created for this (ephemeral) box.
set -u
OUT=${1:-gpurun_out/reference_opencl}
mkdir -p "$OUT" /etc/OpenCL/vendors
[ -f /etc/OpenCL/vendors/nvidia.icd ] || echo "libnvidia-opencl.so.1" > /etc/OpenCL/vendors/nvidia.icd
export LD_LIBRARY_PATH=$PWD/oracle/_ref/clproxy:/usr/local/cuda/targets/x86_64-linux/lib:/usr/local/cuda/lib64:${LD_LIBRARY_PATH:-}
export PTB_REF_CLPROXY_VERBOSE=1
export PTB_REF_WORKDIR=/tmp/ptb_ref_run
rm -rf $PTB_REF_WORKDIR
timeout 600 oracle/_ref/adlTest64 --gtest_filter=DeviceTest.deviceInfo:DeviceTest.RayCast > "$OUT/reference_run.log" 2>&1
echo "reference rc=$?"; sort "$OUT/reference_run.log" | uniq -c | sort -rn | head -14 | cut -c1-200
ls -la $PTB_REF_WORKDIR/build/ | head
cp $PTB_REF_WORKDIR/build/*.ppm "$OUT/reference.ppm" 2>/dev/null
( time oclpathtracer_b200/host/ptb_raycast data/cornellbox.bin "$OUT/ours.ppm" 512 10000 ) 2>&1 | tail -5
python - "$OUT" <<'PY'
import sys, numpy as np, json
out = sys.argv[1]
def load(p):
    t = open(p).read().split()
    assert t[0] == "P3"
    w, h = int(t[1]), int(t[2])
    return np.array(t[4:], np.int32).reshape(h, w, 3)
try:
    a, b = load(out + "/reference.ppm"), load(out + "/ours.ppm")
    d = np.abs(a - b)
    res = {"pixels": int(a.shape[0] * a.shape[1]), "identical_pixels": int((d.max(2) == 0).sum()), "max_abs_diff_8bit": int(d.max()),
           "mean_abs_diff_8bit": float(d.mean()), "pixels_diff_gt_2": int((d.max(2) > 2).sum()),
           "rrmse": float(np.sqrt(((a - b) ** 2).mean()) / np.sqrt((a.astype(float) ** 2).mean()))}
    print(json.dumps(res))
    json.dump(res, open(out + "/compare.json", "w"), indent=1)
except Exception as e:
    print("compare failed:", e)
PY
# the same unmodified test sources compiled against the ADL-shaped shim (oracle/_ref/adlTest64_ptb200 -> libptb200.so)
if [ -x oracle/_ref/adlTest64_ptb200 ]; then
  export PTB_REF_WORKDIR=/tmp/ptb_ref_run_ptb
  rm -rf $PTB_REF_WORKDIR
  timeout 300 oracle/_ref/adlTest64_ptb200 --gtest_filter='DeviceTest.*' > "$OUT/reference_test_on_ptb200.log" 2>&1
  echo "unmodified test on libptb200 rc=$?"; tail -6 "$OUT/reference_test_on_ptb200.log"
  cp $PTB_REF_WORKDIR/build/*.ppm "$OUT/reference_test_on_ptb200.ppm" 2>/dev/null
  cmp "$OUT/reference_test_on_ptb200.ppm" "$OUT/ours.ppm" && echo "PPM identical to host/ptb_raycast"
fi
