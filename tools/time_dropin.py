"""Times the drop-in progressive loop (ptb_launch1d + sync per frame, RaytraceTest.cpp:248-262) on a B200:
FRAME_AHEAD off/on and batch-size multipliers.  Usage: python tools/time_dropin.py [frames] [dim]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oclpathtracer_b200 as pt

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
tris, mats = pt.load_model("data/cornellbox.bin")
dev = pt.Device(0)
tb = dev.buffer(tris.nbytes); mb = dev.buffer(mats.nbytes); fb = dev.buffer(dim * dim * 16)
tb.write(tris); mb.write(mats)
k = dev.kernel("GenerateColors", "GenerateColors")


def loop(n):
    fb.clear()
    dev.sync()
    t0 = time.perf_counter()
    for f in range(n):
        dev.launch1d(k, [tb, mb, fb], pt.Int4(dim, dim, f, 0), dim * dim)
        dev.sync()
    return time.perf_counter() - t0


loop(64)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for ahead, mult in ((0, 0), (1, 0), (1, 1), (1, 2), (1, 3)) * reps:
    dev.kernel_set_int(k, "FRAME_AHEAD", ahead)
    dev.set_tuning(1, mult)
    t = loop(frames)
    print(f"FRAME_AHEAD={ahead} slots=4Mi<<{mult}: {frames} frames of {dim}x{dim} in {t:.3f} s = {t / frames * 1e3:.4f} ms/frame", flush=True)
# the batched entry point on the same work, for reference
scene = dev.scene(tris, mats)
fb.clear(); dev.sync()
t0 = time.perf_counter()
dev.render(scene, pt.default_params(width=dim, height=dim, first_frame=0, n_frames=frames, mode=pt.MODE_PATH, accum=pt.ACCUM_REFERENCE, max_depth=16), fb)
dev.sync()
t = time.perf_counter() - t0
print(f"ptb_render n_frames={frames}: {t:.3f} s = {t / frames * 1e3:.4f} ms/frame")
for accum, name in ((pt.ACCUM_REFERENCE, "reference"), (pt.ACCUM_LINEAR, "linear")):
    for fpb in (0, 4, 1):
        dev.profile(True); dev.profile_read()
        fb.clear(); dev.sync()
        t0 = time.perf_counter()
        dev.render(scene, pt.default_params(width=dim, height=dim, first_frame=0, n_frames=frames, mode=pt.MODE_PATH, accum=accum, max_depth=16, frames_per_batch=fpb), fb)
        dev.sync()
        t = time.perf_counter() - t0
        print(f"ptb_render accum={name} fpb={fpb}: {t:.3f} s; profile {dev.profile_read()}")
        dev.profile(False)
