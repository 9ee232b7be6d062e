"""GPU parity tests: the sm_100a path, called through the C-ABI, against the CPU oracle.

Bit-exact bar: hit triangle ids, t/u/v bits, BVH node-visit counts, ray counts and radiance bits all
equal the oracle's.  The north-star tolerance for radiance (relative RMSE <= 1e-3 at equal spp and RNG
stream) is asserted too -- it is implied by bit equality, and is the bar if a libm-free kernel ever
has to change.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCENE, assert_same_tree, bits, oracle_params, oracle_tree, oracle_tree_for_mode

pytestmark = pytest.mark.gpu

RRMSE_TOL = 1e-3  # north_star: relative RMSE <= 1e-3 at equal spp with an identical RNG stream


def rrmse(a, b):
    a = a[:, :3].astype(np.float64)
    b = b[:, :3].astype(np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30))


@pytest.fixture(scope="module")
def scene(dev, cornell):
    tris, mats = cornell
    s = dev.scene(tris, mats)
    yield s
    s.close()


# ---- numerics contract ---------------------------------------------------------------------------

def test_device_is_blackwell(dev):
    assert "sm_10" in dev.name() and dev.sm_count() >= 100


def test_sincos_bit_exact(dev, ob):
    x = np.concatenate([np.linspace(0, 2 * np.pi, 2_000_001), np.random.default_rng(0).uniform(0, 6.2831855, 500_000),
                        [0.0, 6.2831855, 1.5707964, 3.1415927, 4.712389]]).astype(np.float32)
    s, c = dev.test_sincos(x)
    so, co = ob.sincos(x)
    np.testing.assert_array_equal(bits(s), bits(so))
    np.testing.assert_array_equal(bits(c), bits(co))


def test_pow_bit_exact(dev, ob):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 1, 400_000), rng.uniform(0, 300, 400_000), 10.0 ** rng.uniform(-44, 38, 100_000),
                        [0.0, 1.0, np.inf, np.nan, -1.0, 1e-45, 3.4e38]]).astype(np.float32)
    for y in (np.float32(2.2), np.float32(1.0) / np.float32(2.2)):
        np.testing.assert_array_equal(bits(dev.test_pow(x, y)), bits(ob.powf(x, y)))


def test_rng_known_answers_on_device(dev):
    kat = json.load(open(os.path.join(GOLDEN, "rng_kat.json")))
    for case in kat["cases"]:
        st, va = dev.test_rng(case["gid"], case["frame"], len(case["states"]))
        assert st.tolist() == case["states"]
        assert bits(va).tolist() == case["value_bits"]


@pytest.mark.parametrize("w,h,frame", [(512, 512, 0), (1920, 1080, 7), (100, 37, 3)])
def test_camera_bit_exact(dev, ob, w, h, frame):
    rng = np.random.default_rng(w)
    gids = np.unique(np.concatenate([rng.integers(0, w * h, 2000), [0, w - 1, w * h - 1]])).astype(np.int32)
    o, d, s = dev.test_camera(w, h, frame, gids)
    for k, gid in enumerate(gids):
        seed = (int(gid) + ob.lib().ora_hash_uint32(frame)) & 0xFFFFFFFF
        oo, od, os_ = ob.generate_ray(int(gid) % w, int(gid) // w, w, h, seed)
        assert bits(o[k]).tolist() == bits(oo).tolist() and bits(d[k]).tolist() == bits(od).tolist() and s[k] == os_


# ---- scene queries ------------------------------------------------------------------------------------

def _rays(n, seed):
    rng = np.random.default_rng(seed)
    o = rng.uniform([-2.7, 0.05, -5.5], [2.7, 5.4, 3.9], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    d[:64] = np.float32([0, 0, -1])
    d[64:128] = np.float32([0, -1, 0])
    d[128:192] = np.float32([-1, 0, 0])
    d[192:200] = np.float32([np.nan, 0, 1])  # NaN rays must terminate identically
    return o, d


@pytest.mark.parametrize("any_hit", [False, True])
def test_trace_hit_ids_and_visit_counts(dev, pt, ob, cornell, cornell_bvh, scene, any_hit):
    tris, _ = cornell
    _, bvh, _ = cornell_bvh
    o, d = _rays(300_000, 11)
    tmax = np.float32(2.5) if any_hit else np.float32(1e20)
    g_bvh = dev.trace(scene, o, d, tmax, accel=pt.ACCEL_BVH, any_hit=any_hit)
    g_bru = dev.trace(scene, o, d, tmax, accel=pt.ACCEL_BRUTE, any_hit=any_hit)
    o_bvh = ob.trace(tris, o, d, tmax, bvh=bvh, any_hit=any_hit)
    o_bru = ob.trace(tris, o, d, tmax, bvh=None, any_hit=any_hit)
    for f in ("tri", "t", "u", "v", "visits", "tests"):  # same traversal, same counts
        np.testing.assert_array_equal(bits(g_bvh[f]), bits(o_bvh[f]), err_msg=f"bvh {f}")
        np.testing.assert_array_equal(bits(g_bru[f]), bits(o_bru[f]), err_msg=f"brute {f}")
    if not any_hit:  # closest hit is order independent: BVH == the reference's brute-force loop
        for f in ("tri", "t", "u", "v"):
            np.testing.assert_array_equal(bits(g_bvh[f]), bits(o_bru[f]))
    else:
        np.testing.assert_array_equal(g_bvh["tri"] >= 0, o_bru["tri"] >= 0)
    assert (g_bvh["tests"] <= 36).all() and g_bvh["visits"].max() <= 18  # FLAT form: visits = leaf boxes passed


# ---- the hot path: every mode x integrator x accel ----------------------------------------------------

MODES = {"primary": 0, "ao": 1, "direct": 2, "path": 3}


@pytest.mark.parametrize("integrator", ["mega", "wavefront"])
@pytest.mark.parametrize("accel", ["bvh", "brute"])
@pytest.mark.parametrize("mode", list(MODES))
def test_render_bit_exact(dev, pt, ob, cornell, scene, mode, accel, integrator):
    tris, mats = cornell
    bvh, _keep, ot = oracle_tree_for_mode(ob, tris, MODES[mode])
    assert scene.mode_width(MODES[mode]) == ot["width"] == (4 if mode == "direct" else 1)
    w, h, nf = 96, 80, 3
    use_bvh = accel == "bvh"
    prm = pt.default_params(width=w, height=h, n_frames=nf, mode=MODES[mode], accum=pt.ACCUM_LINEAR, max_depth=8,
                            accel=pt.ACCEL_BVH if use_bvh else pt.ACCEL_BRUTE, collect_stats=1, frames_per_batch=2,
                            integrator=pt.INTEGRATOR_MEGAKERNEL if integrator == "mega" else pt.INTEGRATOR_WAVEFRONT)
    frame = dev.buffer(w * h * 16)
    stats = dev.buffer(w * h * 32)
    ctr = dev.render(scene, prm, frame, stats, want_counters=True)
    fb = frame.read(np.float32).reshape(-1, 4)
    st = stats.read(pt.STATS_DTYPE)
    frame.close(); stats.close()
    oprm = oracle_params(ob, tris, w, h, n_frames=nf, mode=MODES[mode], accum=ob.ACCUM_LINEAR, max_depth=8,
                         use_bvh=1 if use_bvh else 0)
    ofb, ost, octr = ob.render(oprm, tris, mats, bvh=bvh if use_bvh else None, want_stats=True)
    for f in ost.dtype.names:
        np.testing.assert_array_equal(st[f], ost[f], err_msg=f)
    np.testing.assert_array_equal(bits(fb), bits(ofb))
    assert rrmse(fb, ofb) <= RRMSE_TOL
    for k in ("rays_closest", "rays_any", "nodes", "tri_tests", "samples"):
        assert ctr[k] == octr[k], k


@pytest.mark.parametrize("width", [1, 4, 2])
def test_golden_vectors_on_device(dev, pt, cornell, width):
    """tests/golden/oracle_small.npz: the oracle's own trees and 32x32 renders in the three scene forms
    (FLAT = what ptb_scene_create picks for the Cornell box, 4-wide and binary forced)."""
    g = np.load(os.path.join(GOLDEN, "oracle_small.npz"))
    tris, mats = cornell
    sc = dev.scene(tris, mats, pt.bvh_params(force_width=0 if width == 1 else width))
    sfx = {1: "", 4: "_w4", 2: "_w2"}[width]
    assert sc.info()["width"] == width
    nodes, order = sc.bvh()
    assert nodes.view(np.uint8).tobytes() == g["bvh_nodes" + sfx].tobytes() and order.tolist() == g["bvh_order" + sfx].tolist()
    if width == 2:  # the quantised encoding derived on the device equals the committed fixture
        q, lo, step = sc.bvh_quantized()
        assert q.tobytes() == g["bvh_qnodes_w2"].tobytes() and np.array(lo + step, np.float32).tobytes() == g["bvh_qgrid_w2"].tobytes()
    for name, mode in MODES.items():
        sfx = {1: "", 4: "_w4", 2: "_w2"}[sc.mode_width(mode)]  # a FLAT scene walks its 4-wide tree for DIRECT
        for integ in (pt.INTEGRATOR_MEGAKERNEL, pt.INTEGRATOR_WAVEFRONT):
            prm = pt.default_params(width=32, height=32, n_frames=3, mode=mode, accum=pt.ACCUM_LINEAR, max_depth=8, collect_stats=1,
                                    integrator=integ)
            frame, stats = dev.buffer(32 * 32 * 16), dev.buffer(32 * 32 * 32)
            ctr = dev.render(sc, prm, frame, stats, want_counters=True)
            fb, st = frame.read(np.float32), stats.read(pt.STATS_DTYPE)
            frame.close(); stats.close()
            assert fb.tobytes() == g[f"{name}_fb{sfx}"].tobytes(), (name, integ)
            assert st.tobytes() == g[f"{name}_stats{sfx}"].tobytes(), (name, integ)
            assert [ctr[k] for k in ("rays_closest", "rays_any", "nodes", "tri_tests")] == g[f"{name}_ctr{sfx}"].tolist()
    sc.close()


@pytest.mark.parametrize("width", [4, 2])
@pytest.mark.parametrize("any_hit", [False, True])
def test_trace_forced_tree_forms(dev, pt, ob, cornell, width, any_hit):
    """the 4-wide and binary while-while traversals on the Cornell box (force_width), against the oracle walking its own tree"""
    tris, mats = cornell
    sc = dev.scene(tris, mats, pt.bvh_params(force_width=width))
    bvh, _keep, ot = oracle_tree(ob, tris, width=width)
    assert_same_tree(*sc.bvh(), ot)
    o, d = _rays(100_000, 12)
    tmax = np.float32(2.5) if any_hit else np.float32(1e20)
    g = dev.trace(sc, o, d, tmax, accel=pt.ACCEL_BVH, any_hit=any_hit)
    r = ob.trace(tris, o, d, tmax, bvh=bvh, any_hit=any_hit)
    for f in ("tri", "t", "u", "v", "visits", "tests"):
        np.testing.assert_array_equal(bits(g[f]), bits(r[f]), err_msg=f)
    sc.close()


# ---- the reference's own flow: GenerateColors launched frame by frame ---------------------------------

def test_launch1d_drop_in_flow(dev, pt, ob, cornell):
    """RaytraceTest.cpp:216-268: upload tBuffer/materialBuffer, launch GenerateColors per frame with
    int4{W,H,frame,-}, read the gamma-space framebuffer back."""
    tris, mats = cornell
    w = h = 64
    nf = 6
    tb = dev.buffer(36 * 64 * 64)   # the reference over-allocates: nElems is passed a byte count (:222)
    mb = dev.buffer(18 * 64 * 64)
    fbuf = dev.buffer(w * h * 16)
    tb.write(tris); mb.write(mats)
    k = dev.kernel("../test/ClKernels/GenerateColors", "GenerateColors")
    for frame in range(nf):
        dev.launch1d(k, [tb, mb, fbuf], pt.Int4(w, h, frame, 0), w * h)
    dev.sync()
    got = fbuf.read(np.float32).reshape(-1, 4)
    want, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=nf, mode=3, accum=ob.ACCUM_REFERENCE, use_bvh=0), tris, mats)
    np.testing.assert_array_equal(bits(got), bits(want))
    assert (pt.to_rgb8(got) == ob.to_rgb8(want)).all()
    # brute-force variant of the same kernel (ACCEL option) gives the same bits
    dev.kernel_set_int(k, "ACCEL", pt.ACCEL_BRUTE)
    fbuf.clear()
    for frame in range(nf):
        dev.launch1d(k, [tb, mb, fbuf], pt.Int4(w, h, frame, 0), w * h)
    dev.sync()
    np.testing.assert_array_equal(bits(fbuf.read(np.float32).reshape(-1, 4)), bits(want))
    dev.kernel_set_int(k, "ACCEL", pt.ACCEL_BVH)
    # changing the scene buffer is picked up (scene cache keyed on buffer version)
    t2 = tris.copy(); t2["p1"][10:12, 1] -= 1.0; t2["p2"][10:12, 1] -= 1.0; t2["p3"][10:12, 1] -= 1.0
    tb.write(t2); fbuf.clear()
    dev.launch1d(k, [tb, mb, fbuf], pt.Int4(w, h, 0, 0), w * h)
    dev.sync()
    want2, _, _ = ob.render(ob.default_params(w, h, n_frames=1, mode=3, accum=ob.ACCUM_REFERENCE), t2, mats)
    np.testing.assert_array_equal(bits(fbuf.read(np.float32).reshape(-1, 4)), bits(want2))
    with pytest.raises(pt.PtbError, match="n_threads"):
        dev.launch1d(k, [tb, mb, fbuf], pt.Int4(w, h, 0, 0), w * h + 64)
    with pytest.raises(pt.PtbError, match="no kernel"):
        dev.kernel("../test/ClKernels/TestKernel", "VectorAdd")
    for b in (tb, mb, fbuf):
        b.close()


def test_reference_accum_batched_equals_per_frame(dev, pt, ob, cornell, scene):
    tris, mats = cornell
    w, h = 48, 48
    want, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=7, mode=3, accum=ob.ACCUM_REFERENCE), tris, mats)
    for integ in (pt.INTEGRATOR_MEGAKERNEL, pt.INTEGRATOR_WAVEFRONT):
        frame = dev.buffer(w * h * 16)
        dev.render(scene, pt.default_params(width=w, height=h, first_frame=0, n_frames=3, frames_per_batch=2, integrator=integ), frame)
        dev.render(scene, pt.default_params(width=w, height=h, first_frame=3, n_frames=4, frames_per_batch=3, integrator=integ), frame)
        np.testing.assert_array_equal(bits(frame.read(np.float32).reshape(-1, 4)), bits(want))
        frame.close()


# ---- sharding, ragged sizes, degenerate scenes ----------------------------------------------------------

def test_sharded_render_is_bit_identical(dev, pt, ob, cornell, scene):
    from oclpathtracer_b200 import sharding
    import torch

    tris, mats = cornell
    w, h, world, block = 100, 37, 4, 64  # 3700 pixels: ragged
    full, _, _ = ob.render(ob.default_params(w, h, n_frames=2, mode=3, accum=1, max_depth=6), tris, mats)
    parts = []
    for r in range(world):
        prm = pt.default_params(width=w, height=h, n_frames=2, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=6,
                                shard_index=r, shard_count=world, shard_block=block)
        fb, _, _ = dev.render_host(tris, mats, prm)
        assert len(fb) == sharding.local_pixels(w * h, r, world, block)
        parts.append(torch.from_numpy(fb))
    img = sharding.assemble(parts, w * h, world, block).numpy()
    np.testing.assert_array_equal(bits(img), bits(full))


def test_single_triangle_and_miss_paths(dev, pt, ob, cornell):
    tris, mats = cornell
    one = tris[10:11].copy()  # half of the light, seen from below
    p1, ea, eb = pt.light_from_quad(tris, 5)
    for mode in (0, 1, 2, 3):
        prm = pt.default_params(width=64, height=64, n_frames=2, mode=mode, accum=pt.ACCUM_LINEAR, max_depth=4, collect_stats=1,
                                light_p1=p1, light_ea=ea, light_eb=eb)
        for integ in (pt.INTEGRATOR_MEGAKERNEL, pt.INTEGRATOR_WAVEFRONT):
            prm.integrator = integ
            fb, st, ctr = dev.render_host(one, mats, prm, want_stats=True)
            bvh, _keep, _ = oracle_tree_for_mode(ob, one, mode)
            oprm = oracle_params(ob, tris, 64, 64, n_frames=2, mode=mode, accum=1, max_depth=4, use_bvh=1)
            ofb, ost, octr = ob.render(oprm, one, mats, bvh=bvh, want_stats=True)
            np.testing.assert_array_equal(bits(fb), bits(ofb))
            assert st.tobytes() == ost.tobytes()
            assert (st["tri"] == -1).any() and (st["tri"] == 0).any()


@pytest.mark.parametrize("stack", ["local", "shared"])
def test_tessellated_scene_global_memory_path(dev, pt, ob, cornell, stack):
    """k=12 -> 5184 triangles: nodes beyond the shared-memory prefix and triangles come from global memory;
    the traversal stack lives in local memory (default) or in shared memory (tune[2]=2)."""
    dev.set_tuning(2, 2 if stack == "shared" else 0)
    try:
        _tessellated_body(dev, pt, ob, cornell)
    finally:
        for k in (2, 4, 13):
            dev.set_tuning(k, 0)


def _tessellated_body(dev, pt, ob, cornell):
    tris, mats = cornell
    big = pt.tessellate(tris, 12)
    p1, ea, eb = pt.light_from_quad(tris, 5)
    bp = pt.bvh_params(smem_nodes=128)
    sc = dev.scene(big, mats, bp)
    info = sc.info()
    assert info["n_nodes"] > info["smem_nodes"] == 128 and info["width"] == 2
    dev.set_tuning(4, 128)  # stage the whole 128-node prefix (default cap is 64)
    nodes, order = sc.bvh()
    bvh, _keep, ot = oracle_tree(ob, big, width=2, smem_nodes=128)  # the oracle's own tree; the product's must be the same
    assert_same_tree(nodes, order, ot)
    w, h = 64, 64
    for mode in (0, 1, 2, 3):
        for integ, t13 in ((pt.INTEGRATOR_MEGAKERNEL, 0), (pt.INTEGRATOR_WAVEFRONT, 0), (pt.INTEGRATOR_WAVEFRONT, 1)):  # tune[13] = 1: state-machine extend
            dev.set_tuning(13, t13)
            prm = pt.default_params(width=w, height=h, n_frames=2, mode=mode, accum=pt.ACCUM_LINEAR, max_depth=6,
                                    collect_stats=1, integrator=integ, light_p1=p1, light_ea=ea, light_eb=eb)
            frame, stats = dev.buffer(w * h * 16), dev.buffer(w * h * 32)
            ctr = dev.render(sc, prm, frame, stats, want_counters=True)
            fb, st = frame.read(np.float32).reshape(-1, 4), stats.read(pt.STATS_DTYPE)
            frame.close(); stats.close()
            oprm = ob.default_params(w, h, n_frames=2, mode=mode, accum=1, max_depth=6, use_bvh=1, light_p1=p1, light_ea=ea, light_eb=eb)
            ofb, ost, octr = ob.render(oprm, big, mats, bvh=bvh, want_stats=True)
            np.testing.assert_array_equal(bits(fb), bits(ofb), err_msg=f"mode {mode} integ {integ}")
            assert st.tobytes() == ost.tobytes()
            assert ctr["nodes"] == octr["nodes"] and ctr["tri_tests"] == octr["tri_tests"]
    # sampled ground truth: BVH hits == the reference's brute-force loop over all 5184 triangles
    o, d = _rays(20_000, 5)
    g = dev.trace(sc, o, d, np.float32(1e20), accel=pt.ACCEL_BVH)
    r = ob.trace(big, o, d, np.float32(1e20), bvh=None)
    for f in ("tri", "t", "u", "v"):
        np.testing.assert_array_equal(bits(g[f]), bits(r[f]))
    dev.set_tuning(4, 0)
    dev.set_tuning(13, 0)
    sc.close()


# ---- BASELINE.json configurations at full size: size-independent properties -----------------------------

def _render(dev, pt, scene, **kw):
    prm = pt.default_params(**kw)
    n = pt.local_pixels(prm)
    frame = dev.buffer(n * 16)
    ctr = dev.render(scene, prm, frame, None, want_counters=True)
    fb = frame.read(np.float32).reshape(-1, 4)
    frame.close()
    return fb, ctr


def test_c1_primary_512_full_size_vs_oracle(dev, pt, ob, cornell, cornell_bvh, scene):
    """configs[0]: 512x512 primary rays, 1 spp, hit-ID output -- small enough to check every pixel."""
    tris, mats = cornell
    _, bvh, _ = cornell_bvh
    prm = pt.default_params(width=512, height=512, n_frames=1, mode=pt.MODE_PRIMARY, accum=pt.ACCUM_LINEAR, collect_stats=1)
    fb, st, ctr = dev.render_host(tris, mats, prm, want_stats=True)
    ofb, ost, octr = ob.render(ob.default_params(512, 512, mode=0, accum=1, use_bvh=1), tris, mats, bvh=bvh, want_stats=True)
    brute, bst, _ = ob.render(ob.default_params(512, 512, mode=0, accum=1, use_bvh=0), tris, mats, want_stats=True)
    assert st.tobytes() == ost.tobytes() and fb.tobytes() == ofb.tobytes()
    np.testing.assert_array_equal(st["tri"], bst["tri"])  # == the reference's brute-force hit ids
    np.testing.assert_array_equal(st["t_bits"], bst["t_bits"])
    assert ctr["rays_closest"] == 512 * 512


def _full_size_vs_oracle(dev, pt, ob, cornell, scene, w, h, mode, frames, shard=None, **kw):
    """One full-size frame range of a BASELINE configuration, rendered through the C-ABI, against the ORACLE walking its own
    tree: radiance bits, per-pixel hit ids / t bits / node-visit counts / id hashes, and the ray / node / test counters."""
    tris, mats = cornell
    bvh, _keep, ot = oracle_tree_for_mode(ob, tris, mode)
    assert scene.mode_width(mode) == ot["width"]
    p1, ea, eb = pt.light_from_quad(tris, 5)
    skw = dict(shard_index=shard[0], shard_count=shard[1], shard_block=shard[2]) if shard else {}
    for first, n in frames:
        prm = pt.default_params(width=w, height=h, first_frame=first, n_frames=n, mode=mode, accum=pt.ACCUM_LINEAR, collect_stats=1,
                                light_p1=p1, light_ea=ea, light_eb=eb, **skw, **kw)
        nl = pt.local_pixels(prm)
        frame, stats = dev.buffer(nl * 16), dev.buffer(nl * 32)
        ctr = dev.render(scene, prm, frame, stats, want_counters=True)
        fb, st = frame.read(np.float32).reshape(-1, 4), stats.read(pt.STATS_DTYPE)
        frame.close(); stats.close()
        oprm = ob.default_params(w, h, first_frame=first, n_frames=n, mode=mode, accum=ob.ACCUM_LINEAR, use_bvh=1,
                                 light_p1=p1, light_ea=ea, light_eb=eb, **skw, **kw)
        ofb, ost, octr = ob.render(oprm, tris, mats, bvh=bvh, want_stats=True)
        np.testing.assert_array_equal(bits(fb), bits(ofb), err_msg=f"radiance, frames {first}+{n}")
        for f in ost.dtype.names:
            np.testing.assert_array_equal(st[f], ost[f], err_msg=f"{f}, frames {first}+{n}")
        for k in ("rays_closest", "rays_any", "nodes", "tri_tests", "samples"):
            assert ctr[k] == octr[k], (k, first, n)
    return fb, ctr


def test_c2_ao_1024_full_size_vs_oracle(dev, pt, ob, cornell, scene):
    """configs[1] at FULL size: AO 1024x1024, 16 rays/pixel -- the whole frame against the oracle (17.8 Mrays), then the
    size-independent properties: BVH == brute force == wavefront bit for bit, values are k/16."""
    a, ca = _full_size_vs_oracle(dev, pt, ob, cornell, scene, 1024, 1024, pt.MODE_AO, [(0, 1)], ao_samples=16)
    kw = dict(width=1024, height=1024, n_frames=1, mode=pt.MODE_AO, accum=pt.ACCUM_LINEAR, ao_samples=16)
    b, cb = _render(dev, pt, scene, accel=pt.ACCEL_BRUTE, integrator=pt.INTEGRATOR_MEGAKERNEL, **kw)
    c, cc = _render(dev, pt, scene, accel=pt.ACCEL_BVH, integrator=pt.INTEGRATOR_WAVEFRONT, **kw)
    assert a.tobytes() == b.tobytes() == c.tobytes()
    assert ca["rays_closest"] == 1024 * 1024 and ca["rays_any"] == 16 * 1024 * 1024 == cc["rays_any"] == cb["rays_any"]
    v = a[:, 0] * 16
    assert np.array_equal(v, np.round(v)) and v.min() >= 0 and v.max() <= 16
    assert np.array_equal(a[:, 0], a[:, 1]) and np.array_equal(a[:, 3], np.ones(len(a), np.float32))


def test_c3_direct_1080p_full_size_vs_oracle(dev, pt, ob, cornell, scene):
    """configs[2] at FULL size: direct lighting 1920x1080 -- frames 0-1 and the LAST frame of the 64-spp range (63) against
    the oracle, every pixel; then integrators and accelerators agree bit for bit on 4 frames."""
    _full_size_vs_oracle(dev, pt, ob, cornell, scene, 1920, 1080, pt.MODE_DIRECT, [(0, 2), (63, 1)])
    kw = dict(width=1920, height=1080, n_frames=4, mode=pt.MODE_DIRECT, accum=pt.ACCUM_LINEAR)
    a, ca = _render(dev, pt, scene, accel=pt.ACCEL_BVH, integrator=pt.INTEGRATOR_MEGAKERNEL, **kw)
    b, cb = _render(dev, pt, scene, accel=pt.ACCEL_BRUTE, integrator=pt.INTEGRATOR_WAVEFRONT, **kw)
    assert a.tobytes() == b.tobytes() and ca["rays_any"] == cb["rays_any"] > 0
    assert np.isfinite(a).all() and a[:, :3].min() >= 0


def test_c4_path_4k_full_size_vs_oracle(dev, pt, ob, cornell, scene):
    """configs[3] at FULL size: 3840x2160, max depth 8 -- one whole 4K frame (frame 0: 8.3 M paths, ~30 M rays) and the LAST
    frame of the 256-spp range (255, a 1/4 tile shard) against the oracle; then megakernel == wavefront and linearity of
    the frame mean: mean(f0,f1) == (f0 + f1) / 2 from single-frame renders."""
    _full_size_vs_oracle(dev, pt, ob, cornell, scene, 3840, 2160, pt.MODE_PATH, [(0, 1)], max_depth=8)
    _full_size_vs_oracle(dev, pt, ob, cornell, scene, 3840, 2160, pt.MODE_PATH, [(255, 1)], shard=(1, 4, 64), max_depth=8)
    kw = dict(width=3840, height=2160, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=8)
    a, ca = _render(dev, pt, scene, n_frames=2, integrator=pt.INTEGRATOR_MEGAKERNEL, **kw)
    b, cb = _render(dev, pt, scene, n_frames=2, integrator=pt.INTEGRATOR_WAVEFRONT, **kw)
    assert a.tobytes() == b.tobytes() and ca["rays_closest"] == cb["rays_closest"]
    f0, _ = _render(dev, pt, scene, first_frame=0, n_frames=1, **kw)
    f1, _ = _render(dev, pt, scene, first_frame=1, n_frames=1, **kw)
    want = (f0[:, :3] + f1[:, :3]) / np.float32(2.0)
    np.testing.assert_array_equal(bits(a[:, :3]), bits(want))
    assert ca["rays_closest"] <= 8 * 2 * 3840 * 2160


def test_buffer_map_unmap_and_errors(dev, pt):
    import ctypes as C
    b = dev.buffer(1024)
    hp = C.c_void_p()
    assert pt.lib().ptb_buffer_map(b._h, C.byref(hp)) == 0
    dev.sync()
    arr = (C.c_uint8 * 1024).from_address(hp.value)
    for i in range(1024):
        arr[i] = i & 255
    assert pt.lib().ptb_buffer_unmap(b._h, hp) == 0
    dev.sync()
    assert b.read(np.uint8).tolist() == [i & 255 for i in range(1024)]
    assert pt.lib().ptb_buffer_unmap(b._h, C.c_void_p(1234)) != 0
    with pytest.raises(pt.PtbError, match="range exceeds"):
        b.write(np.zeros(2048, np.uint8))
    b.close()
    tiny = dev.buffer(16)
    with pytest.raises(pt.PtbError, match="too small"):
        sc_tris, sc_mats = pt.load_model(SCENE)
        s = dev.scene(sc_tris, sc_mats)
        try:
            dev.render(s, pt.default_params(width=8, height=8), tiny)
        finally:
            s.close()
    tiny.close()


def test_cpp_raycast_flow_matches_oracle(ob, cornell, tmp_path):
    """host/raycast_main.cpp = the reference's DeviceTest.RayCast flow over the ADL-shaped shim."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oclpathtracer_b200", "host", "ptb_raycast")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    out = tmp_path / "rc.ppm"
    r = subprocess.run([exe, SCENE, str(out), "64", "5"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    vals = [int(v) for v in out.read_text().split()[4:]]
    tris, mats = cornell
    want, _, _ = ob.render(ob.default_params(64, 64, first_frame=0, n_frames=5, mode=3, accum=ob.ACCUM_REFERENCE), tris, mats)
    assert vals == ob.to_rgb8(want).reshape(-1).tolist()


def test_async_host_pipeline(dev, pt, cornell):
    """ptb_render_host_async / ptb_job_wait: double-buffered jobs return exactly the synchronous results."""
    tris, mats = cornell
    w, h, n = 160, 96, 6
    want = [dev.render_host(tris, mats, pt.default_params(width=w, height=h, first_frame=f, n_frames=1, mode=pt.MODE_AO,
                                                          accum=pt.ACCUM_LINEAR), want_counters=False)[0].copy() for f in range(n)]
    pinned = [pt.PinnedArray((w * h, 4), np.float32) for _ in range(2)]
    pageable = [np.zeros((w * h, 4), np.float32) for _ in range(2)]
    for bufs in ([p.array for p in pinned], pageable):
        prev, got = None, []
        for f in range(n):
            prm = pt.default_params(width=w, height=h, first_frame=f, n_frames=1, mode=pt.MODE_AO, accum=pt.ACCUM_LINEAR)
            job = dev.render_host_async(tris, mats, prm, bufs[f & 1])
            if prev is not None:
                dev.job_wait(prev[0])
                got.append(bufs[prev[1] & 1].copy())
            prev = (job, f)
        dev.job_wait(prev[0])
        got.append(bufs[prev[1] & 1].copy())
        for f in range(n):
            assert got[f].tobytes() == want[f].tobytes(), f
    # a third job in flight is refused, and the synchronous call refuses to run under in-flight jobs
    prm = pt.default_params(width=w, height=h, mode=pt.MODE_AO, accum=pt.ACCUM_LINEAR)
    j1 = dev.render_host_async(tris, mats, prm, pageable[0])
    j2 = dev.render_host_async(tris, mats, prm, pageable[1])
    with pytest.raises(pt.PtbError, match="two jobs"):
        dev.render_host_async(tris, mats, prm, pinned[0].array)
    with pytest.raises(pt.PtbError, match="in flight"):
        dev.render_host(tris, mats, prm)
    # waiting for the NEWER job first frees its slot for the next submit (any free slot is taken, not jobs_submitted & 1)
    dev.job_wait(j2)
    j3 = dev.render_host_async(tris, mats, prm, pageable[1])
    dev.job_wait(j1); dev.job_wait(j3)
    for p in pinned:
        p.free()


def test_async_reference_accumulation_is_sequential(dev, pt, ob, cornell):
    """accum=REFERENCE continues the gamma-space running mean from the caller's out_rgba, read at submit time: a pipelined
    submit while another job is in flight would use a stale state, so it is refused; the wait-then-submit loop equals the
    synchronous path and the oracle, with pinned and with pageable buffers."""
    tris, mats = cornell
    w, h, n = 96, 64, 7
    want, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=n, mode=3, accum=ob.ACCUM_REFERENCE, max_depth=6), tris, mats)
    pin = pt.PinnedArray((w * h, 4), np.float32)
    for buf in (pin.array, np.zeros((w * h, 4), np.float32)):
        buf[...] = 0
        for f in range(n):
            prm = pt.default_params(width=w, height=h, first_frame=f, n_frames=1, mode=pt.MODE_PATH, accum=pt.ACCUM_REFERENCE, max_depth=6)
            dev.job_wait(dev.render_host_async(tris, mats, prm, buf))
        np.testing.assert_array_equal(bits(buf), bits(want))
    other = np.zeros((w * h, 4), np.float32)
    j = dev.render_host_async(tris, mats, pt.default_params(width=w, height=h, first_frame=0, n_frames=1, mode=pt.MODE_PATH, max_depth=6), other)
    with pytest.raises(pt.PtbError, match="wait for the job in flight"):
        dev.render_host_async(tris, mats, pt.default_params(width=w, height=h, first_frame=1, n_frames=1, mode=pt.MODE_PATH, max_depth=6), pin.array)
    dev.job_wait(j)
    pin.free()


def test_launch1d_sees_buffers_marked_dirty(dev, pt, cornell):
    """A tBuffer rewritten behind its back (here: through a second view of its device pointer) keeps rendering the old scene until
    ptb_buffer_mark_dirty tells ptb_launch1d to re-read it (the reference kernel re-reads its buffers on every launch)."""
    tris, mats = cornell
    w, h = 64, 48
    tb, mb, fb = dev.buffer(36 * 64), dev.buffer(18 * 64), dev.buffer(w * h * 16)
    tb.write(tris); mb.write(mats)
    k = dev.kernel("GenerateColors", "GenerateColors")
    dev.kernel_set_int(k, "BOUNCES", 16)
    dev.launch1d(k, [tb, mb, fb], pt.Int4(w, h, 0, 0), w * h); dev.sync()
    first = fb.read(np.uint32).copy()
    t2 = tris.copy()
    for key in ("p1", "p2", "p3"):
        t2[key][16:36, 1] += 1.0  # lift both boxes
    alias = dev.wrap(tb.device_ptr(), 36 * 64)
    alias.write(t2)
    alias.close()
    dev.launch1d(k, [tb, mb, fb], pt.Int4(w, h, 0, 0), w * h); dev.sync()
    assert np.array_equal(fb.read(np.uint32), first)  # documented: the resident scene is keyed on API writes
    tb.mark_dirty()
    dev.launch1d(k, [tb, mb, fb], pt.Int4(w, h, 0, 0), w * h); dev.sync()
    lifted = fb.read(np.uint32).copy()
    assert not np.array_equal(lifted, first)
    tb2 = dev.buffer(36 * 64); tb2.write(t2)
    dev.launch1d(k, [tb2, mb, fb], pt.Int4(w, h, 0, 0), w * h); dev.sync()
    assert np.array_equal(fb.read(np.uint32), lifted)
    for b in (tb, tb2, mb, fb):
        b.close()


def test_path_regeneration_equals_one_sample_per_thread(dev, pt, scene):
    """tune[5]: persistent path megakernel with regeneration (default) vs the plain one-sample-per-thread form."""
    kw = dict(width=333, height=77, n_frames=3, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=9, collect_stats=1,
              integrator=pt.INTEGRATOR_MEGAKERNEL, frames_per_batch=2)
    res = []
    try:
        for variant in (0, 1):
            dev.set_tuning(5, variant)
            frame, stats = dev.buffer(333 * 77 * 16), dev.buffer(333 * 77 * 32)
            ctr = dev.render(scene, pt.default_params(**kw), frame, stats, want_counters=True)
            res.append((frame.read(np.uint32).tobytes(), stats.read(np.uint32).tobytes(), tuple(sorted(ctr.items()))))
            frame.close(); stats.close()
    finally:
        dev.set_tuning(5, 0)
    assert res[0] == res[1]


@pytest.mark.parametrize("build", ["host_cornell", "host_tess", "device_lbvh"])
def test_quantised_nodes_equal_the_oracles(dev, pt, ob, cornell, build):
    """Binary trees are traversed through their 32-byte quantised encoding (ptb_bvh_nodeq), derived on the device from the fp32
    nodes for host-built and device-built trees alike: its bytes and grid equal the oracle's restatement (ora_bvh_quantize)
    applied to the same fp32 nodes."""
    tris, mats = cornell
    if build == "host_cornell":
        sc = dev.scene(tris, mats, pt.bvh_params(force_width=2))
    elif build == "host_tess":
        sc = dev.scene(pt.tessellate(tris, 20), mats)
    else:
        sc = dev.scene(pt.tessellate(tris, 9), mats, gpu_build=True)
    assert sc.info()["width"] == 2
    nodes, _order = sc.bvh()
    q, lo, step = sc.bvh_quantized()
    oq, olo, ostep = ob.quantize(nodes)
    assert lo == olo and step == ostep
    assert q.tobytes() == oq.tobytes()
    sc.close()


@pytest.mark.parametrize("form", ["flat", "wide4", "binary_smem", "binary_global"])
def test_path_state_machine_is_scheduling_only(dev, pt, cornell, form):
    """k_path_sm (per-lane state machine, the warp votes which phase runs: tune[5]=3) against k_mega_path_regen (tune[5]=2)
    and the one-sample-per-thread megakernel (tune[5]=1), for every scene form and for quorum thresholds from 1 to 32
    (tune[0] regen, tune[10] shade, tune[11] leaf): radiance bits, per-pixel statistics and counters must not move."""
    tris, mats = cornell
    if form == "binary_global":
        sc = dev.scene(pt.tessellate(tris, 12), mats)
        assert sc.info()["width"] == 2 and sc.info()["n_nodes"] > 64
    else:
        sc = dev.scene(tris, mats, pt.bvh_params(force_width={"flat": 1, "wide4": 4, "binary_smem": 2}[form]))
    w, h = 333, 77
    kw = dict(width=w, height=h, n_frames=3, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=9, collect_stats=1,
              integrator=pt.INTEGRATOR_MEGAKERNEL, frames_per_batch=2)
    res = []
    try:
        # tune[12]: 0 = k_path_sm2 (path state parked in shared memory; large scenes on the local-memory stack), 1 = k_path_sm, 28 / 29 / 30 = sm2 at 8 / 9 / 10 CTAs (default 11)
        for variant, thr, t12 in ((1, (0, 0, 0), 0), (2, (0, 0, 0), 0), (3, (0, 0, 0), 0), (3, (1, 1, 1), 0), (3, (32, 32, 32), 0), (3, (3, 20, 2), 0),
                                  (3, (12, 2, 27), 0), (3, (0, 0, 0), 1), (3, (2, 31, 3), 1), (3, (0, 0, 0), 28), (3, (0, 0, 0), 29), (3, (0, 0, 0), 30), (0, (0, 0, 0), 0), (0, (1, 1, 1), 0)):
            dev.set_tuning(5, variant)
            dev.set_tuning(12, t12)
            for k, v in zip((0, 10, 11), thr):
                dev.set_tuning(k, v)
            for stats_on in (1, 0):
                frame, stats = dev.buffer(w * h * 16), dev.buffer(w * h * 32)
                ctr = dev.render(sc, pt.default_params(**{**kw, "collect_stats": stats_on}), frame, stats if stats_on else None, want_counters=True)
                res.append(((variant, t12), thr, stats_on, frame.read(np.uint32).tobytes(), stats.read(np.uint32).tobytes() if stats_on else b"",
                            (ctr["rays_closest"], ctr["samples"]) + ((ctr["nodes"], ctr["tri_tests"]) if stats_on else ())))
                frame.close(); stats.close()
    finally:
        for k in (5, 0, 10, 11, 12):
            dev.set_tuning(k, 0)
        sc.close()
    for r in res:
        ref = res[0] if r[2] else res[1]
        assert r[3] == ref[3], (form, r[:3], "radiance")
        assert r[4] == ref[4], (form, r[:3], "stats")
        assert r[5] == ref[5], (form, r[:3], "counters")


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_wavefront_persistent_fetch_is_scheduling_only(dev, pt, scene, mode):
    """tune[6]/tune[7]: persistent traversal with dynamic ray fetch (any refill threshold) vs one thread per ray."""
    kw = dict(width=250, height=90, n_frames=3, mode=mode, accum=pt.ACCUM_LINEAR, max_depth=7, collect_stats=1,
              integrator=pt.INTEGRATOR_WAVEFRONT, frames_per_batch=2, ao_samples=5)
    res = []
    try:
        for persist_off, thr in ((1, 0), (0, 1), (0, 8), (0, 20), (0, 32)):
            dev.set_tuning(6, persist_off); dev.set_tuning(7, thr)
            frame, stats = dev.buffer(250 * 90 * 16), dev.buffer(250 * 90 * 32)
            ctr = dev.render(scene, pt.default_params(**kw), frame, stats, want_counters=True)
            res.append((frame.read(np.uint32).tobytes(), stats.read(np.uint32).tobytes(), tuple(sorted(ctr.items()))))
            frame.close(); stats.close()
    finally:
        dev.set_tuning(6, 0); dev.set_tuning(7, 0)
    for r in res[1:]:
        assert r == res[0]


def test_launch_capture_and_replay(dev, pt, ob, cornell, tmp_path):
    """Launcher::serializeToFile / deserializeFromFile analogue, in the reference's own file layout."""
    import struct
    tris, mats = cornell
    w = h = 48
    tb, mb, fb = dev.buffer(36 * 64), dev.buffer(18 * 64), dev.buffer(w * h * 16)
    tb.write(tris); mb.write(mats)
    k = dev.kernel("GenerateColors", "GenerateColors")
    for frame in range(3):
        dev.launch1d(k, [tb, mb, fb], pt.Int4(w, h, frame, 0), w * h)
    dev.sync()
    dump = str(tmp_path / "launch.bin")
    dev.launch_serialize(dump, [tb, mb, fb], pt.Int4(w, h, 3, 0), w * h)   # state BEFORE frame 3
    dev.launch1d(k, [tb, mb, fb], pt.Int4(w, h, 3, 0), w * h)
    dev.sync()
    direct = fb.read(np.uint32)
    raw = open(dump, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == 4 and struct.unpack_from("<ii", raw, 4) == (1, 36 * 64)
    assert struct.unpack_from("<7i", raw, len(raw) - 28) == (w * h, 1, 1, 64, 1, 1, 1)
    bufs, consts, n_threads, local = dev.launch_deserialize(dump)
    assert len(bufs) == 3 and n_threads == w * h and local == 64 and struct.unpack("<4i", consts) == (w, h, 3, 0)
    c = pt.Int4(*struct.unpack("<4i", consts))
    dev.launch1d(k, bufs, c, n_threads, local)
    dev.sync()
    assert bufs[2].read(np.uint32).tobytes() == direct.tobytes()
    want, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=4, mode=3, accum=ob.ACCUM_REFERENCE), tris, mats)
    assert direct.tobytes() == want.tobytes()
    with pytest.raises(pt.PtbError, match="cannot open"):
        dev.launch_deserialize(str(tmp_path / "nope.bin"))
    (tmp_path / "short.bin").write_bytes(raw[:100])
    with pytest.raises(pt.PtbError, match="truncated"):
        dev.launch_deserialize(str(tmp_path / "short.bin"))
    for b in bufs + [tb, mb, fb]:
        b.close()


def _check_tree(nodes, order, tris, pad_min):
    """every triangle reachable exactly once from the root; every child box contains its triangles (+pad)."""
    n = len(tris)
    W = 4 if nodes.dtype.itemsize == 128 else 2
    tri_lo = np.minimum(np.minimum(tris["p1"], tris["p2"]), tris["p3"])[:, :3]
    tri_hi = np.maximum(np.maximum(tris["p1"], tris["p2"]), tris["p3"])[:, :3]
    covered = np.zeros(n, np.int32)
    stack = [(0, None, None)]
    depth_max = 0
    work = [(0, 1)]
    bounds = {}
    # iterative post-order: compute subtree bounds
    order_stack, visit = [0], []
    while order_stack:
        ni = order_stack.pop()
        visit.append(ni)
        for ch in (int(nodes[f"child{k}"][ni]) for k in range(W)):
            if ch >= 0 and ch != 0x7FFFFFFF:
                order_stack.append(ch)
    assert len(set(visit)) == len(visit)

    def leaf_bounds(ref):
        code = (~ref) & 0xFFFFFFFF
        first, count = code >> 3, (code & 7) + 1
        assert 1 <= count <= 8 and first + count <= n
        covered[first:first + count] += 1
        idx = order[first:first + count]
        return tri_lo[idx].min(0), tri_hi[idx].max(0)

    for ni in reversed(visit):
        nd = nodes[ni]
        res = []
        for k in range(W):
            ch, c, e = int(nd[f"child{k}"]), nd[f"c{k}"].astype(np.float64), nd[f"e{k}"].astype(np.float64)
            if ch == 0x7FFFFFFF:
                assert (e < 0).all()
                continue
            l, h = leaf_bounds(ch) if ch < 0 else bounds[ch]
            assert (c - e <= l - pad_min).all() and (c + e >= h + pad_min).all()
            res.append((l, h))
        bounds[ni] = (np.minimum.reduce([r[0] for r in res]), np.maximum.reduce([r[1] for r in res]))
    assert (covered == 1).all()
    assert sorted(order.tolist()) == list(range(n))


@pytest.mark.parametrize("k,max_leaf", [(1, 4), (6, 2), (20, 4)])
def test_gpu_lbvh_builder(dev, pt, ob, cornell, k, max_leaf):
    """BVH built on the device: valid tree, hits == the reference's brute-force loop, renders == oracle on that tree."""
    tris, mats = cornell
    scene_tris = tris if k == 1 else pt.tessellate(tris, k)
    p1, ea, eb = pt.light_from_quad(tris, 5)
    sc = dev.scene(scene_tris, mats, pt.bvh_params(max_leaf=max_leaf), gpu_build=True)
    info = sc.info()
    assert info["n_nodes"] == len(scene_tris) - 1 and info["smem_nodes"] == 1 and 1 <= info["depth"] <= 64 and info["width"] == 2
    nodes, order = sc.bvh()
    _check_tree(nodes, order, scene_tris, pad_min=5e-4)
    # deterministic: a second build gives the same bytes
    sc2 = dev.scene(scene_tris, mats, pt.bvh_params(max_leaf=max_leaf), gpu_build=True)
    n2, o2 = sc2.bvh()
    assert np.array_equal(o2, order)
    live = np.zeros(len(nodes), bool)
    st = [0]
    while st:
        i = st.pop(); live[i] = True
        st += [int(nodes[f"child{k}"][i]) for k in range(2) if 0 <= nodes[f"child{k}"][i] != 0x7FFFFFFF]
    assert nodes[live].tobytes() == n2[live].tobytes()
    sc2.close()
    o, d = _rays(30_000, 17)
    g = dev.trace(sc, o, d, np.float32(1e20), accel=pt.ACCEL_BVH)
    r = ob.trace(scene_tris, o, d, np.float32(1e20), bvh=None)
    for f in ("tri", "t", "u", "v"):
        np.testing.assert_array_equal(bits(g[f]), bits(r[f]))
    bvh, _keep = ob.make_bvh(nodes, order)
    gv = ob.trace(scene_tris, o, d, np.float32(1e20), bvh=bvh)
    np.testing.assert_array_equal(g["visits"], gv["visits"])
    w, h = 72, 56
    for mode, integ in ((1, pt.INTEGRATOR_MEGAKERNEL), (3, pt.INTEGRATOR_MEGAKERNEL), (3, pt.INTEGRATOR_WAVEFRONT), (2, pt.INTEGRATOR_WAVEFRONT)):
        prm = pt.default_params(width=w, height=h, n_frames=2, mode=mode, accum=pt.ACCUM_LINEAR, max_depth=5, collect_stats=1,
                                integrator=integ, light_p1=p1, light_ea=ea, light_eb=eb)
        frame, stats = dev.buffer(w * h * 16), dev.buffer(w * h * 32)
        ctr = dev.render(sc, prm, frame, stats, want_counters=True)
        fb, stt = frame.read(np.float32).reshape(-1, 4), stats.read(pt.STATS_DTYPE)
        frame.close(); stats.close()
        oprm = ob.default_params(w, h, n_frames=2, mode=mode, accum=1, max_depth=5, use_bvh=1, light_p1=p1, light_ea=ea, light_eb=eb)
        ofb, ost, octr = ob.render(oprm, scene_tris, mats, bvh=bvh, want_stats=True)
        np.testing.assert_array_equal(bits(fb), bits(ofb))
        assert stt.tobytes() == ost.tobytes() and ctr["nodes"] == octr["nodes"]
    sc.close()


def test_reference_device_tests_over_shim():
    """host/adl_tests.cpp: the reference's DeviceTest cases (test/main.cpp:53-152 + RayCast) over the ADL-shaped shim."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oclpathtracer_b200", "host", "adl_tests")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    r = subprocess.run([exe, SCENE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("[       OK ]") == 7


def test_c5_two_million_triangles_properties(dev, pt, ob, cornell):
    """configs[4] at FULL size: the 2,005,056-triangle scene (18 quads x 236^2 x 2) at 3840x2160, max depth 8.
    A 1/16 tile shard of frame 0 and a 1/64 shard of the LAST frame of the 64-spp range (63) equal the CPU oracle walking
    its own tree (which the product's tree must equal byte for byte): radiance bits, hit ids, node-visit counts, counters.
    Size-independent properties: radiance does not depend on the acceleration structure (host SAH tree == device LBVH
    tree == megakernel == wavefront, bit for bit); sampled hit ids equal the reference's brute-force loop over all 2M triangles."""
    tris, mats = cornell
    big = pt.tessellate(tris, 236)
    assert len(big) == 2_005_056
    p1, ea, eb = pt.light_from_quad(tris, 5)
    w, h = 3840, 2160
    kw = dict(width=w, height=h, n_frames=1, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=8,
              light_p1=p1, light_ea=ea, light_eb=eb)
    frames = {}
    sah = dev.scene(big, mats)
    assert sah.info()["width"] == 2 and sah.info()["n_nodes"] > 500_000
    for name, integ in (("sah_mega", pt.INTEGRATOR_MEGAKERNEL), ("sah_wave", pt.INTEGRATOR_WAVEFRONT)):
        frame = dev.buffer(w * h * 16)
        ctr = dev.render(sah, pt.default_params(integrator=integ, **kw), frame, None, want_counters=True)
        frames[name] = (frame.read(np.uint32).tobytes(), ctr["rays_closest"])
        frame.close()
    # sampled ground truth: the reference's brute-force loop (on the GPU: 2M Moller-Trumbore tests per ray)
    o, d = _rays(1500, 23)
    g = dev.trace(sah, o, d, np.float32(1e20), accel=pt.ACCEL_BVH)
    r = dev.trace(sah, o, d, np.float32(1e20), accel=pt.ACCEL_BRUTE)
    for f in ("tri", "t", "u", "v"):
        np.testing.assert_array_equal(bits(g[f]), bits(r[f]))
    assert (r["tests"] == len(big)).all() and g["tests"].max() < 200
    # a shard of the image against the CPU oracle walking its own tree (which the product's must equal)
    nodes, order = sah.bvh()
    bvh, _keep, ot = oracle_tree(ob, big)
    assert_same_tree(nodes, order, ot)
    for first, shard in ((0, dict(shard_index=5, shard_count=16, shard_block=64)), (63, dict(shard_index=40, shard_count=64, shard_block=64))):
        prm = pt.default_params(**{**kw, "first_frame": first}, collect_stats=1, **shard)
        nl = pt.local_pixels(prm)
        fbuf, sbuf = dev.buffer(nl * 16), dev.buffer(nl * 32)
        ctr = dev.render(sah, prm, fbuf, sbuf, want_counters=True)
        got, gst = fbuf.read(np.float32).reshape(-1, 4), sbuf.read(pt.STATS_DTYPE)
        fbuf.close(); sbuf.close()
        okw = dict(first_frame=first, n_frames=1, mode=3, accum=1, max_depth=8, use_bvh=1, light_p1=p1, light_ea=ea, light_eb=eb, **shard)
        want, wst, octr = ob.render(ob.default_params(w, h, **okw), big, mats, bvh=bvh, want_stats=True)
        np.testing.assert_array_equal(bits(got), bits(want))
        assert gst.tobytes() == wst.tobytes()
        for k in ("rays_closest", "nodes", "tri_tests", "samples"):
            assert ctr[k] == octr[k], (k, first)
    sah.close()
    lb = dev.scene(big, mats, gpu_build=True)
    frame = dev.buffer(w * h * 16)
    ctr = dev.render(lb, pt.default_params(integrator=pt.INTEGRATOR_MEGAKERNEL, **kw), frame, None, want_counters=True)
    frames["lbvh_mega"] = (frame.read(np.uint32).tobytes(), ctr["rays_closest"])
    frame.close(); lb.close()
    assert frames["sah_mega"] == frames["sah_wave"] == frames["lbvh_mega"]


def test_against_the_genuine_reference_output(dev, pt, cornell, scene):
    """The reference's own DeviceTest.RayCast (10000 frames, 512x512, gamma-space running mean, sqrt + 8-bit PPM),
    run UNMODIFIED on a B200 through NVIDIA's OpenCL (tools/run_reference_opencl.sh; tests/golden/
    reference_raycast_b200_opencl.npz), against the same flow through libptb200.so.  OpenCL's sin/cos/pow/normalize
    are implementation-defined, so this is a tolerance test: north-star rRMSE <= 1e-3 at equal spp and RNG stream
    (measured 9.95e-4 on the 8-bit output, where one-level rounding flips dominate)."""
    g = np.load(os.path.join(GOLDEN, "reference_raycast_b200_opencl.npz"))["rgb"].astype(np.int32)
    w = h = 512
    frame = dev.buffer(w * h * 16)
    dev.render(scene, pt.default_params(width=w, height=h, first_frame=0, n_frames=10000, mode=pt.MODE_PATH,
                                        accum=pt.ACCUM_REFERENCE, max_depth=16), frame)
    ours = pt.to_rgb8(frame.read(np.float32).reshape(-1, 4)).reshape(h, w, 3).astype(np.int32)
    frame.close()
    d = np.abs(ours - g)
    rr = float(np.sqrt(((ours - g) ** 2).mean()) / np.sqrt((g.astype(np.float64) ** 2).mean()))
    assert rr <= 1.5e-3, rr
    assert (d.max(2) == 0).mean() >= 0.97
    assert (d.max(2) > 2).sum() <= 60


def test_unmodified_reference_test_program_on_libptb200(dev, pt, cornell, scene, tmp_path):
    """oracle/_ref/adlTest64_ptb200 = the reference's own test/RaytraceTest.cpp + test/main.cpp, compiled UNMODIFIED with
    oclpathtracer_b200/host first on the include path (make -C oracle ref; needs /root/reference, so the binary is
    prebuilt and travels).  Its DeviceTest.RayCast (10000 launches + waitForCompletion, then its own PPM writer) must
    (1) pass, (2) write exactly the image the batched C-ABI render gives, (3) agree with the image the same program
    wrote through the reference's OpenCL backend on this GPU (tests/golden/reference_raycast_b200_opencl.npz)."""
    import glob
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "adlTest64_ptb200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/adlTest64_ptb200 not built (make -C oracle ref needs the reference tree)")
    env = dict(os.environ, PTB_REF_WORKDIR=str(tmp_path / "ref"))
    r = subprocess.run([exe, "--gtest_filter=DeviceTest.*"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "[  PASSED  ] 6 tests" in r.stdout and "FAILED" not in r.stdout, r.stdout[-1500:]  # all its DeviceTest cases
    ppm = glob.glob(str(tmp_path / "ref" / "build" / "rayCastAo_*.ppm"))
    assert len(ppm) == 1, ppm
    t = open(ppm[0]).read().split()
    assert t[:4] == ["P3", "512", "512", "255"]
    got = np.array(t[4:], np.int32).reshape(512, 512, 3)
    frame = dev.buffer(512 * 512 * 16)
    dev.render(scene, pt.default_params(width=512, height=512, first_frame=0, n_frames=10000, mode=pt.MODE_PATH,
                                        accum=pt.ACCUM_REFERENCE, max_depth=16), frame)
    ours = pt.to_rgb8(frame.read(np.float32).reshape(-1, 4)).reshape(512, 512, 3).astype(np.int32)
    frame.close()
    assert np.array_equal(got, ours)
    g = np.load(os.path.join(GOLDEN, "reference_raycast_b200_opencl.npz"))["rgb"].astype(np.int32)
    rr = float(np.sqrt(((got - g) ** 2).mean()) / np.sqrt((g.astype(np.float64) ** 2).mean()))
    assert rr <= 1.5e-3, rr


def test_launch1d_frame_ahead_batching(dev, pt, ob, cornell):
    """ptb_launch1d traces the next frames of a consecutive-frame loop together (frame-ahead batching) and folds one
    frame per launch.  The framebuffer after EVERY launch must equal the one-launch-per-frame path bit for bit, also
    across a scene rewrite in mid-sequence, a jump in the frame index, a repeated frame, a caller-side clear of the
    framebuffer and a change of BOUNCES."""
    tris, mats = cornell
    w, h = 96, 64
    tb = dev.buffer(36 * 64); mb = dev.buffer(18 * 64)
    tb.write(tris); mb.write(mats)
    k = dev.kernel("GenerateColors", "GenerateColors")
    t2 = tris.copy()
    for key in ("p1", "p2", "p3"):
        t2[key][10:12, 1] -= 0.75
    # (frame index, action before the launch)
    script = [(f, None) for f in range(0, 23)] + [(23, "rewrite")] + [(f, None) for f in range(24, 40)] + \
             [(100, None), (101, None), (102, None), (102, None), (103, "clear"), (104, None), (105, "bounces"),
              (106, None), (107, None), (3, None), (4, None), (5, None)]

    def run(frame_ahead):
        dev.kernel_set_int(k, "FRAME_AHEAD", frame_ahead)
        dev.kernel_set_int(k, "BOUNCES", 16)
        tb.write(tris)
        fb = dev.buffer(w * h * 16)
        states = []
        for frame, action in script:
            if action == "rewrite":
                tb.write(t2)
            elif action == "clear":
                fb.clear()
            elif action == "bounces":
                dev.kernel_set_int(k, "BOUNCES", 5)
            dev.launch1d(k, [tb, mb, fb], pt.Int4(w, h, frame, 0), w * h)
            dev.sync()
            states.append(bits(fb.read(np.float32)).copy())
        fb.close()
        return states

    plain = run(0)
    ahead = run(1)
    for i, (a, b) in enumerate(zip(plain, ahead)):
        assert np.array_equal(a, b), f"launch {i} (frame {script[i][0]}) differs"
    # and the plain path is the oracle's, through the first scene rewrite
    want, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=23, mode=3, accum=ob.ACCUM_REFERENCE), tris, mats)
    assert np.array_equal(plain[22], bits(want).reshape(-1))
    dev.kernel_set_int(k, "FRAME_AHEAD", 1)
    dev.kernel_set_int(k, "BOUNCES", 16)
    tb.close(); mb.close()


# ---- several GPUs in one process: ptb_device_add_helper / ptb_render_multi -------------------------------------------------

def _helper_indices(pt, n):
    """device indices for n helpers: the other GPUs of the box when there are any, else device 0 again (a second handle on the same
    GPU takes every code path -- sharding, shared tree, cross-stream events -- only the stores are not remote)"""
    import ctypes as C
    c = C.c_int(0)
    pt.lib().ptb_device_count(C.byref(c))
    return [(1 + i) % max(1, c.value) for i in range(n)], c.value


@pytest.mark.parametrize("n_helpers", [1, 3])
def test_render_multi_is_bit_identical(pt, ob, cornell, n_helpers):
    """ptb_render_multi: ONE image tile-sharded over a device and its helpers, host records in, host image out ==
    ptb_render_host on one device == the oracle, for every mode, both accumulators and the 8-bit output."""
    tris, mats = cornell
    idx, n_gpus = _helper_indices(pt, n_helpers)
    main = pt.Device(0)
    single = pt.Device(0)
    helpers = [pt.Device(i) for i in idx]
    try:
        for hlp in helpers:
            main.add_helper(hlp)
        assert main.helper_count() == n_helpers
        with pytest.raises(pt.PtbError, match="serves one device"):
            single.add_helper(helpers[0])
        w, h = 200, 131
        for mode in (0, 1, 2, 3):
            prm = pt.default_params(width=w, height=h, n_frames=3, mode=mode, accum=pt.ACCUM_LINEAR, max_depth=6, frames_per_batch=2)
            got, ctr = main.render_multi(tris, mats, prm)
            want, _, wctr = single.render_host(tris, mats, prm)
            np.testing.assert_array_equal(bits(got), bits(want), err_msg=f"mode {mode}")
            assert ctr["rays_closest"] == wctr["rays_closest"] and ctr["rays_any"] == wctr["rays_any"] and ctr["samples"] == w * h * 3
        # the reference's progressive state (gamma-space running mean) carried across calls, then the 8-bit output
        fb = np.zeros((w * h, 4), np.float32)
        for f0, nf in ((0, 2), (2, 3), (5, 1)):
            prm = pt.default_params(width=w, height=h, first_frame=f0, n_frames=nf, mode=pt.MODE_PATH, accum=pt.ACCUM_REFERENCE, max_depth=5)
            main.render_multi(tris, mats, prm, out=fb, want_counters=False)
        ofb, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=6, mode=3, accum=ob.ACCUM_REFERENCE, max_depth=5), tris, mats)
        np.testing.assert_array_equal(bits(fb), bits(ofb))
        prm = pt.default_params(width=w, height=h, n_frames=4, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=5, output=pt.OUTPUT_RGB8)
        rgb, _ = main.render_multi(tris, mats, prm)
        lin, _, _ = single.render_host(tris, mats, pt.default_params(width=w, height=h, n_frames=4, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=5))
        np.testing.assert_array_equal(rgb, ob.to_rgb8(lin))
        # a scene traversed from L2/HBM: every device gets the shared host-built tree and derives its own quantised nodes
        big = pt.tessellate(tris, 12)
        prm = pt.default_params(width=w, height=h, n_frames=2, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=6)
        got, ctr = main.render_multi(big, mats, prm)
        want, _, wctr = single.render_host(big, mats, prm)
        np.testing.assert_array_equal(bits(got), bits(want), err_msg="tessellated scene")
        assert ctr["rays_closest"] == wctr["rays_closest"]
        # a changed scene is picked up by every device; an image with fewer blocks than devices still renders
        t2 = tris.copy()
        t2["p1"][10:12, 1] -= 0.5
        prm = pt.default_params(width=64, height=2, n_frames=2, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR, max_depth=4)
        a, _ = main.render_multi(t2, mats, prm)
        b, _, _ = single.render_host(t2, mats, prm)
        np.testing.assert_array_equal(bits(a), bits(b))
    finally:
        helpers[0].close()   # a helper may go first ...
        main.close()         # ... or after its device
        for hlp in helpers[1:]:
            hlp.close()
        single.close()


def test_launch1d_deals_frame_ahead_batches_over_helpers(pt, ob, cornell):
    """The reference's progressive loop (RaytraceTest.cpp:250-268) through ptb_launch1d on a device WITH helpers: the frames of
    each frame-ahead batch are traced by all devices into the main device's batch; the framebuffer after every launch equals
    the single-device loop and, at the end, the oracle's."""
    tris, mats = cornell
    idx, n_gpus = _helper_indices(pt, 2)
    main, single = pt.Device(0), pt.Device(0)
    helpers = [pt.Device(i) for i in idx]
    try:
        for hlp in helpers:
            main.add_helper(hlp)
        w, h, nf = 128, 64, 45
        states = {}
        for name, d in (("multi", main), ("single", single)):
            tb, mb, fb = d.buffer(36 * 64), d.buffer(18 * 64), d.buffer(w * h * 16)
            tb.write(tris); mb.write(mats)
            k = d.kernel("../test/ClKernels/GenerateColors", "GenerateColors")
            out = []
            for frame in range(nf):
                if frame == 30:  # the scene changes in mid-sequence: every device must drop its copy
                    t2 = tris.copy()
                    t2["p1"][10:12, 1] -= 0.75
                    tb.write(t2)
                d.launch1d(k, [tb, mb, fb], pt.Int4(w, h, frame, 0), w * h)
                d.sync()
                out.append(fb.read(np.uint32).copy())
            states[name] = out
            for b in (tb, mb, fb):
                b.close()
        for i, (a, b) in enumerate(zip(states["multi"], states["single"])):
            assert np.array_equal(a, b), f"frame {i}"
        want, _, _ = ob.render(ob.default_params(w, h, first_frame=0, n_frames=30, mode=3, accum=ob.ACCUM_REFERENCE), tris, mats)
        assert np.array_equal(states["multi"][29], bits(want).reshape(-1))
    finally:
        main.close()
        for hlp in helpers:
            hlp.close()
        single.close()


@pytest.mark.parametrize("integrator", ["mega", "wavefront"])
def test_ragged_and_degenerate_shapes(dev, pt, ob, cornell, scene, integrator):
    """Edge shapes: a single pixel, sizes below one warp / not a multiple of the CTA, a shard that owns a ragged tail,
    one AO ray, depth 1, a batch size that does not divide the frame count, frame indices near INT_MAX (the seed is
    gid + 1103515245*frame + 12345 in uint32 arithmetic, GenerateColors.cl:305-308).  All bit-exact vs the oracle."""
    tris, mats = cornell
    integ = pt.INTEGRATOR_MEGAKERNEL if integrator == "mega" else pt.INTEGRATOR_WAVEFRONT
    cases = [
        dict(w=1, h=1, mode=3, n_frames=5, fpb=2, first=0, max_depth=16),
        dict(w=7, h=3, mode=1, n_frames=3, fpb=2, first=0, ao_samples=1),
        dict(w=33, h=17, mode=2, n_frames=4, fpb=3, first=2147483000),
        dict(w=130, h=1, mode=3, n_frames=3, fpb=0, first=7, max_depth=1),
        dict(w=61, h=29, mode=3, n_frames=3, fpb=2, first=0, max_depth=4, shard=(2, 3, 50)),
        dict(w=40, h=40, mode=0, n_frames=1, fpb=1, first=0, shard=(6, 7, 64)),
    ]
    for cs in cases:
        bvh, _keep, _ = oracle_tree_for_mode(ob, tris, cs["mode"])
        kw = dict(n_frames=cs["n_frames"], first_frame=cs["first"], mode=cs["mode"], max_depth=cs.get("max_depth", 8),
                  ao_samples=cs.get("ao_samples", 16))
        if "shard" in cs:
            kw.update(shard_index=cs["shard"][0], shard_count=cs["shard"][1], shard_block=cs["shard"][2])
        for accum in (pt.ACCUM_LINEAR, pt.ACCUM_REFERENCE):
            prm = pt.default_params(width=cs["w"], height=cs["h"], accum=accum, frames_per_batch=cs["fpb"],
                                    integrator=integ, collect_stats=1, **kw)
            n_local = pt.local_pixels(prm)
            frame = dev.buffer(max(n_local, 1) * 16)
            stats = dev.buffer(max(n_local, 1) * 32)
            frame.clear()
            ctr = dev.render(scene, prm, frame, stats, want_counters=True)
            fb = frame.read(np.float32).reshape(-1, 4)[:n_local]
            st = stats.read(pt.STATS_DTYPE)[:n_local]
            frame.close(); stats.close()
            oprm = oracle_params(ob, tris, cs["w"], cs["h"], accum=accum, use_bvh=1, **kw)
            ofb, ost, octr = ob.render(oprm, tris, mats, bvh=bvh, want_stats=True)
            assert len(ofb) == n_local, cs
            np.testing.assert_array_equal(bits(fb), bits(ofb), err_msg=str(cs))
            for f in ost.dtype.names:
                np.testing.assert_array_equal(st[f], ost[f], err_msg=f"{f} {cs}")
            for k in ("rays_closest", "rays_any", "nodes", "tri_tests"):
                assert ctr[k] == octr[k], (k, cs)


def test_no_device_memory_leak(pt, cornell):
    """Devices, scenes, buffers, the host pipeline and the frame-ahead batch release what they allocate."""
    tris, mats = cornell
    probe = pt.Device(0)
    free0 = None
    for it in range(6):
        d = pt.Device(0)
        sc = d.scene(tris, mats)
        fb = d.buffer(256 * 256 * 16)
        d.render(sc, pt.default_params(width=256, height=256, n_frames=3, mode=pt.MODE_PATH, accum=pt.ACCUM_LINEAR), fb)
        d.render(sc, pt.default_params(width=256, height=256, n_frames=2, mode=pt.MODE_AO, accum=pt.ACCUM_LINEAR,
                                       integrator=pt.INTEGRATOR_WAVEFRONT), fb)
        d.render_host(tris, mats, pt.default_params(width=128, height=128, n_frames=1, mode=pt.MODE_DIRECT))
        tb = d.buffer(tris.nbytes); mb = d.buffer(mats.nbytes)
        tb.write(tris); mb.write(mats)
        k = d.kernel("GenerateColors", "GenerateColors")
        for f in range(12):  # long enough to allocate a frame-ahead batch
            d.launch1d(k, [tb, mb, fb], pt.Int4(256, 256, f, 0), 256 * 256)
        d.sync()
        for b in (tb, mb, fb):
            b.close()
        sc.close()
        d.close()
        free, _ = probe.memory()
        if it == 1:
            free0 = free  # after the first full cycle (context-level pools are warm)
        elif it > 1:
            assert free >= free0 - (8 << 20), (it, free0, free)
    probe.close()


def test_render_gather_single_process(dev, pt, cornell, scene):
    """ptb_render_gather without peers: four shards rendered one after the other into ONE full image give exactly the
    single-device image (pixels land at their global position; REFERENCE accumulation resumes from that image)."""
    w, h, world, block = 130, 47, 4, 64
    full = dev.buffer(w * h * 16)
    ref = dev.buffer(w * h * 16)
    for accum, chunks in ((pt.ACCUM_LINEAR, [(0, 3)]), (pt.ACCUM_REFERENCE, [(0, 2), (2, 2)])):
        full.clear(); ref.clear()
        for first, n in chunks:
            for r in range(world):
                dev.render_gather(scene, pt.default_params(width=w, height=h, first_frame=first, n_frames=n, mode=pt.MODE_AO, accum=accum,
                                                           shard_index=r, shard_count=world, shard_block=block), full)
            dev.render(scene, pt.default_params(width=w, height=h, first_frame=first, n_frames=n, mode=pt.MODE_AO, accum=accum), ref)
        np.testing.assert_array_equal(full.read(np.uint32), ref.read(np.uint32))
    with pytest.raises(pt.PtbError, match="too small"):
        small = dev.buffer(64)
        try:
            dev.render_gather(scene, pt.default_params(width=w, height=h, shard_index=0, shard_count=2), small)
        finally:
            small.close()
    full.close(); ref.close()


def test_render_gather_over_ipc(pt):
    """Two processes (both on cuda:0 here; one per GPU on a multi-GPU box) map each other's full image with CUDA IPC and
    render their shard with the exchange fused into the resolve kernel; both images equal the single-device render."""
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, PTB_TEST_ONE_GPU="1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "tests", "_gather_worker.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("identical=True") == 4, r.stdout


def test_ieee_helpers_bit_exact(dev):
    """rcp_rn / sqrt_rn / safe_rcp3 / normalize issue the IEEE refinement sequence without the compiler's range dispatch
    where the operand range allows it and the generic operator elsewhere: the results must be the IEEE results (numpy
    float32 division and sqrt are correctly rounded) for every magnitude, sign, zero, denormal, inf and NaN."""
    rng = np.random.default_rng(1234)
    n = 1 << 21
    mag = np.exp2(rng.uniform(-149, 128, n)).astype(np.float32)           # log-uniform over the whole float range
    x = (mag * rng.choice(np.float32([-1, 1]), n)).astype(np.float32)
    edge = np.float32([0.0, -0.0, 1e-30, 1.0000001e-30, 9.999999e-31, 1e30, 9.999999e29, 1.0000001e30, 1e-20, 1.0000001e-20,
                       9.999999e-21, np.inf, -np.inf, np.nan, 1.17549435e-38, 1e-45, 3.4028235e38, 1.0, 2.0, 0.5, 3.0, 1e-8])
    edge = np.concatenate([edge, -edge])
    bits_all = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)  # uniform over bit patterns
    x = np.concatenate([x, edge, bits_all, rng.uniform(0, 1, n).astype(np.float32)])
    got = dev.test_ieee(x)
    with np.errstate(all="ignore"):
        one = np.float32(1.0)
        want_rcp = one / x
        want_sqrt = np.sqrt(x)
        a = np.abs(x)
        want_safe = np.where(a > np.float32(1e-20), one / x, np.where(np.signbit(x), np.float32(-1e20), np.float32(1e20))).astype(np.float32)
        vx, vy, vz = x, np.float32(0.5) * x, np.float32(2.0) * x
        dd = (vx * vx + vy * vy) + vz * vz
        inv = one / np.sqrt(dd)
        want_n = [vx * inv, vy * inv, vz * inv]

    def same(g, w, name):
        g, w = np.asarray(g, np.float32), np.asarray(w, np.float32)
        ok = (g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w))  # NaN payloads are not part of the contract
        assert ok.all(), (name, x[~ok][:5], g[~ok][:5], w[~ok][:5])

    same(got["rcp"], want_rcp, "rcp")
    same(got["sqrt"], want_sqrt, "sqrt")
    same(got["safe_rcp"], want_safe, "safe_rcp")
    for g, w, nme in zip((got["nx"], got["ny"], got["nz"]), want_n, "xyz"):
        same(g, w, "normalize." + nme)


def test_rgb8_output_equals_host_transform(dev, pt, cornell):
    """params.output = RGB8: the host entry points deliver sqrt -> x255 -> truncate -> clamp done on the device
    (RaytraceTest.cpp:78-83,:283); the bytes must equal ptb_to_rgb8 of the float4 frame, also through the async pipeline,
    for pixel counts that are not a multiple of the kernel's four-pixel groups."""
    tris, mats = cornell
    for (w, h), mode, accum in (((97, 53), pt.MODE_PATH, pt.ACCUM_REFERENCE), ((64, 33), pt.MODE_AO, pt.ACCUM_LINEAR),
                                ((130, 1), pt.MODE_DIRECT, pt.ACCUM_LINEAR), ((1, 1), pt.MODE_PRIMARY, pt.ACCUM_LINEAR)):
        kw = dict(width=w, height=h, first_frame=0, n_frames=5, mode=mode, accum=accum, max_depth=6)
        f4, _, _ = dev.render_host(tris, mats, pt.default_params(**kw))
        u8, _, _ = dev.render_host(tris, mats, pt.default_params(output=pt.OUTPUT_RGB8, **kw))
        assert u8.dtype == np.uint8 and u8.shape == (w * h, 3)
        np.testing.assert_array_equal(u8, pt.to_rgb8(f4))
        pin = pt.PinnedArray((w * h, 3), np.uint8)
        job = dev.render_host_async(tris, mats, pt.default_params(output=pt.OUTPUT_RGB8, **kw), pin.array)
        dev.job_wait(job)
        np.testing.assert_array_equal(pin.array, u8)
        pin.free()
    with pytest.raises(pt.PtbError, match="RGB8"):
        dev.render_host(tris, mats, pt.default_params(width=8, height=8, first_frame=3, n_frames=1, accum=pt.ACCUM_REFERENCE,
                                                      output=pt.OUTPUT_RGB8))
