#!/usr/bin/env python
"""A/B sweep over the experiment knobs (ptb_device_set_tuning) on ONE resident scene: builds the scene once, then for every
setting renders `frames` frames of a BASELINE configuration and prints Mrays/s from the integrator's own CUDA-event time
(ptb_device_profile).  Settings are scheduling-only knobs: the first frame's bytes are compared with the first setting's.

  python tools/sweep_tune.py c5 4 "5=2" "" "10=4" "10=12,11=4" ...        # "" = library defaults
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oclpathtracer_b200 as pt  # noqa: E402
from bench import SCENE, WORKLOADS  # noqa: E402


def main():
    name, frames = sys.argv[1], int(sys.argv[2])
    settings = sys.argv[3:] or [""]
    reps = int(os.environ.get("SWEEP_REPS", "2"))
    wl = WORKLOADS[name]
    tris, mats = pt.load_model(SCENE)
    light = pt.light_from_quad(tris, 5)
    if wl.get("tess"):
        tris = pt.tessellate(tris, wl["tess"])
    dev = pt.Device(0)
    bp = None
    if os.environ.get("SWEEP_FORCE_WIDTH") or os.environ.get("SWEEP_MAX_LEAF") or os.environ.get("SWEEP_SAH_TRAVERSE"):
        bp = pt.bvh_params()
        bp.force_width = int(os.environ.get("SWEEP_FORCE_WIDTH", "0"))
        if os.environ.get("SWEEP_MAX_LEAF"):
            bp.max_leaf = int(os.environ["SWEEP_MAX_LEAF"])
        if os.environ.get("SWEEP_SAH_TRAVERSE"):
            bp.traverse_cost = float(os.environ["SWEEP_SAH_TRAVERSE"])
    scene = dev.scene(tris, mats, bp)
    print("scene:", scene.info(), flush=True)
    w, h = wl["width"], wl["height"]
    frame = dev.buffer(w * h * 16)
    ref = None
    for s in settings:
        knobs = [tuple(int(x) for x in kv.split("=")) for kv in s.split(",") if kv]
        for k in range(16):
            dev.set_tuning(k, 0)
        for k, v in knobs:
            dev.set_tuning(k, v)
        integ = {"": pt.INTEGRATOR_AUTO, "mega": pt.INTEGRATOR_MEGAKERNEL, "wavefront": pt.INTEGRATOR_WAVEFRONT}[os.environ.get("SWEEP_INTEGRATOR", "")]
        p = pt.default_params(width=w, height=h, mode=wl["mode"], accum=pt.ACCUM_LINEAR, first_frame=0, n_frames=frames, integrator=integ,
                              frames_per_batch=int(os.environ.get("SWEEP_FPB", "0")))
        if "ao_samples" in wl:
            p.ao_samples = wl["ao_samples"]
        if "max_depth" in wl:
            p.max_depth = wl["max_depth"]
        p.light_p1[:], p.light_ea[:], p.light_eb[:] = light
        dev.render(scene, p, frame)  # warm-up
        dev.sync()
        best = None
        for _ in range(reps):
            dev.counters(cumulative=1, read=False)
            dev.profile(True); dev.profile_read()
            dev.render(scene, p, frame)
            prof = dev.profile_read(); dev.profile(False)
            ctr = dev.counters(cumulative=0, read=True)
            rays = ctr["rays_closest"] + ctr["rays_any"]
            ms = prof["integrator_ms"]
            if best is None or ms < best[0]:
                best = (ms, rays)
        digest = hashlib.sha256(frame.read(np.uint32).tobytes()).hexdigest()[:12]
        if ref is None:
            ref = digest
        print(f"{name} tune[{s or 'defaults'}]: {best[1] / best[0] / 1e3:9.1f} Mrays/s  integrator {best[0] / frames:8.3f} ms/frame  "
              f"{'same bits' if digest == ref else 'BITS DIFFER ' + digest}", flush=True)
    frame.close(); scene.close(); dev.close()


if __name__ == "__main__":
    main()
