"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oclpathtracer_b200 as pt

tris, mats = pt.load_model(os.path.join(ROOT, "data", "cornellbox.bin"))
p1, ea, eb = pt.light_from_quad(tris, 5)
dev = pt.Device(0)
big = pt.tessellate(tris, 6)
for scene_tris, gpu in ((tris, False), (big, False), (big, True)):
    sc = dev.scene(scene_tris, mats, pt.bvh_params(smem_nodes=32), gpu_build=gpu)
    for mode in (0, 1, 2, 3):
        for integ in (pt.INTEGRATOR_MEGAKERNEL, pt.INTEGRATOR_WAVEFRONT):
            for accel in (pt.ACCEL_BVH, pt.ACCEL_BRUTE):
                if accel == pt.ACCEL_BRUTE and len(scene_tris) > 36:
                    continue
                prm = pt.default_params(width=70, height=33, n_frames=3, mode=mode, accum=pt.ACCUM_REFERENCE, max_depth=5, ao_samples=3,
                                        collect_stats=1, integrator=integ, accel=accel, frames_per_batch=2, light_p1=p1, light_ea=ea, light_eb=eb)
                frame, stats = dev.buffer(70 * 33 * 16), dev.buffer(70 * 33 * 32)
                frame.clear()
                dev.render(sc, prm, frame, stats, want_counters=True)
                frame.close(); stats.close()
    rng = np.random.default_rng(0)
    o = rng.uniform(-2, 2, (999, 3)).astype(np.float32); d = rng.normal(size=(999, 3)).astype(np.float32)
    dev.trace(sc, o, d, np.float32(1e20)); dev.trace(sc, o, d, np.float32(2.0), any_hit=True)
    sc.close()
out = np.zeros((70 * 33, 4), np.float32)
j = dev.render_host_async(tris, mats, pt.default_params(width=70, height=33, mode=1, accum=1), out)
dev.job_wait(j)
dev.test_sincos(np.linspace(0, 6, 100).astype(np.float32)); dev.test_pow(np.linspace(0, 6, 100).astype(np.float32), 2.2)
dev.close()
print("sanitize pass done")
