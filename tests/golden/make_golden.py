"""Regenerates the fixtures in tests/golden/.  Run from the repo root:  python tests/golden/make_golden.py

rng_kat.json   -- known answers for the integer RNG, computed HERE in pure Python straight from the
                  arithmetic of the reference source (test/ClKernels/GenerateColors.cl:47-71 and :308);
                  independent of oracle/ and of the CUDA code.  The reference's tests hold no vectors
                  for this path, so these source-derived integers are the pin.
scene.json     -- facts of test/cornellbox.bin decoded per RaytraceTest.cpp:117-143 (counts, sha256,
                  bounding box), computed with numpy only.
oracle_small.npz -- 32x32 renders of every mode by the oracle (regression pin for the oracle itself and
                  a travelling vector for the GPU tests).  Not a reference output: the reference cannot
                  execute here (OpenCL C, no platform).
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
M = 0xFFFFFFFF


def hash_uint32(x):  # GenerateColors.cl:57 (the Wang branch is '#if 0')
    return (1103515245 * x + 12345) & M


def random_float_state(s):  # GenerateColors.cl:63-68
    s = ((s ^ 61) ^ (s >> 16)) & M
    s = (s + (s << 3)) & M
    s = (s ^ (s >> 4)) & M
    s = (s * 0x27D4EB2D) & M
    s = (s ^ (s >> 15)) & M
    s = (1103515245 * s + 12345) & M
    return s


def u32_to_f32_bits(s):  # (float)(uint) round-to-nearest-even, then * 2^-32 (exact power of two)
    f = np.float32(np.uint32(s)) * np.float32(2.3283064365386963e-10)
    return int(np.float32(f).view(np.uint32))


def rng_kat():
    cases = []
    for gid, frame in [(0, 0), (1, 0), (262143, 0), (0, 1), (131328, 7), (8294399, 255), (12345, 9999)]:
        seed = (gid + hash_uint32(frame)) & M  # :308
        states, values = [], []
        s = seed
        for _ in range(8):
            s = random_float_state(s)
            states.append(s)
            values.append(u32_to_f32_bits(s))
        cases.append({"gid": gid, "frame": frame, "seed0": seed, "states": states, "value_bits": values})
    return {"hash_uint32": {str(f): hash_uint32(f) for f in range(8)}, "cases": cases,
            "max_state_value_bits": u32_to_f32_bits(0xFFFFFFFF)}


def scene_facts():
    raw = open(os.path.join(ROOT, "data", "cornellbox.bin"), "rb").read()
    ints = np.frombuffer(raw, "<i4")
    flts = np.frombuffer(raw, "<f4")
    pos = 0
    n_mesh = int(ints[pos]); pos += 1
    meshes, lo, hi = [], np.full(3, np.inf), np.full(3, -np.inf)
    for _ in range(n_mesh):
        nq = int(ints[pos]); pos += 1
        tag = float(flts[pos]); pos += 1
        pos += 4 * nq
        nv = int(ints[pos]); pos += 1
        v = flts[pos:pos + 4 * nv].reshape(nv, 4); pos += 4 * nv
        lo = np.minimum(lo, v[:, :3].min(0)); hi = np.maximum(hi, v[:, :3].max(0))
        meshes.append({"quads": nq, "verts": nv, "tag": tag})
    assert pos * 4 == len(raw)
    return {"sha256": hashlib.sha256(raw).hexdigest(), "bytes": len(raw), "meshes": meshes,
            "aabb_lo": [float(x) for x in lo], "aabb_hi": [float(x) for x in hi]}


def oracle_small():
    from oracle import binding as ob

    tris, mats = ob.load_model(os.path.join(ROOT, "data", "cornellbox.bin"))
    p1, ea, eb = ob.light_from_quad(tris, 5)
    out = {}
    # the oracle's OWN trees (oracle_bvh.c) in the three forms: FLAT (what the product picks for the Cornell box; the
    # un-suffixed keys), 4-wide and binary (force_width = 4 | 2)
    for width, sfx in ((1, ""), (4, "_w4"), (2, "_w2")):
        b = ob.build_bvh(tris, width=width)
        bvh, _keep = ob.make_bvh(b["nodes"], b["tri_order"])
        out["bvh_nodes" + sfx] = b["nodes"].view(np.uint8).reshape(-1, b["nodes"].dtype.itemsize)
        out["bvh_order" + sfx] = b["tri_order"]
        if width == 2:  # the quantised encoding the product traverses (ptb_bvh_nodeq) and its grid
            q, lo, step = ob.quantize(b["nodes"])
            out["bvh_qnodes_w2"] = q
            out["bvh_qgrid_w2"] = np.array(lo + step, np.float32)
        for name, mode in (("primary", 0), ("ao", 1), ("direct", 2), ("path", 3)):
            prm = ob.default_params(32, 32, n_frames=3, mode=mode, accum=ob.ACCUM_LINEAR, use_bvh=1, max_depth=8,
                                    light_p1=p1, light_ea=ea, light_eb=eb)
            fb, st, ctr = ob.render(prm, tris, mats, bvh=bvh, want_stats=True)
            out[f"{name}_fb{sfx}"] = fb
            out[f"{name}_stats{sfx}"] = st.view(np.uint32).reshape(-1, 8)
            out[f"{name}_ctr{sfx}"] = np.array([ctr[k] for k in ("rays_closest", "rays_any", "nodes", "tri_tests")], np.uint64)
    prm = ob.default_params(32, 32, first_frame=0, n_frames=5, mode=3, accum=ob.ACCUM_REFERENCE, use_bvh=0)
    fb, _, _ = ob.render(prm, tris, mats)
    out["path_reference_accum_fb"] = fb
    return out


if __name__ == "__main__":
    json.dump(rng_kat(), open(os.path.join(HERE, "rng_kat.json"), "w"), indent=1)
    json.dump(scene_facts(), open(os.path.join(HERE, "scene.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **oracle_small())
    print("golden fixtures written to", HERE)
