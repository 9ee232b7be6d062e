"""ctypes binding of the CPU oracle (oracle/liboracle_pt.so).

TEST INFRASTRUCTURE.  Import this only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(oclpathtracer_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_pt.so")


class Float4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


TRIANGLE_DTYPE = np.dtype(
    [("p1", "<f4", 4), ("p2", "<f4", 4), ("p3", "<f4", 4), ("id", "<i4"), ("padding", "u1", 12)]
)
MATERIAL_DTYPE = np.dtype(
    [("albedo", "<f4", 4), ("emissive", "<f4", 4), ("roughness", "<f4"), ("type", "<i4"), ("padding", "u1", 24)]
)
NODE_DTYPE = np.dtype(  # binary node, 64 bytes (ptb_bvh_node)
    [
        ("c0", "<f4", 3), ("child0", "<i4"), ("e0", "<f4", 3), ("child1", "<i4"),
        ("c1", "<f4", 3), ("pad0", "<i4"), ("e1", "<f4", 3), ("pad1", "<i4"),
    ]
)
NODE4_DTYPE = np.dtype(  # 4-wide node, 128 bytes (ptb_bvh_node4)
    [
        ("c0", "<f4", 3), ("child0", "<i4"), ("e0", "<f4", 3), ("child1", "<i4"),
        ("c1", "<f4", 3), ("child2", "<i4"), ("e1", "<f4", 3), ("child3", "<i4"),
        ("c2", "<f4", 3), ("pad0", "<i4"), ("e2", "<f4", 3), ("pad1", "<i4"),
        ("c3", "<f4", 3), ("pad2", "<i4"), ("e3", "<f4", 3), ("pad3", "<i4"),
    ]
)
LEAFBOX_DTYPE = np.dtype(  # flat form, 32 bytes (ptb_bvh_leafbox)
    [("c", "<f4", 3), ("mask_lo", "<u4"), ("e", "<f4", 3), ("mask_hi", "<u4")]
)
STATS_DTYPE = np.dtype(
    [
        ("tri", "<i4"), ("quad", "<i4"), ("t_bits", "<u4"), ("visits_primary", "<u4"),
        ("visits_secondary", "<u4"), ("count", "<u4"), ("id_hash", "<u4"), ("tri_tests", "<u4"),
    ]
)
assert TRIANGLE_DTYPE.itemsize == 64 and MATERIAL_DTYPE.itemsize == 64
assert NODE_DTYPE.itemsize == 64 and NODE4_DTYPE.itemsize == 128 and STATS_DTYPE.itemsize == 32


class Bvh(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p), ("n_nodes", C.c_int32),
        ("tri_order", C.c_void_p), ("n_tris", C.c_int32), ("width", C.c_int32),
        ("qnodes", C.c_void_p), ("q_lo", C.c_float * 3), ("q_step", C.c_float * 3),
    ]


class BvhParams(C.Structure):
    _fields_ = [("max_leaf", C.c_int32), ("pad_rel", C.c_float), ("n_bins", C.c_int32), ("smem_nodes", C.c_int32),
                ("traverse_cost", C.c_float)]


class Params(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("first_frame", C.c_int32), ("n_frames", C.c_int32),
        ("mode", C.c_int32), ("accum", C.c_int32), ("use_bvh", C.c_int32),
        ("max_depth", C.c_int32), ("ao_samples", C.c_int32), ("ao_max_dist", C.c_float),
        ("light_quad", C.c_int32),
        ("light_p1", C.c_float * 3), ("light_ea", C.c_float * 3), ("light_eb", C.c_float * 3),
        ("shard_index", C.c_int32), ("shard_count", C.c_int32), ("shard_block", C.c_int32),
        ("n_threads", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "rays_closest", "rays_any", "nodes", "tri_tests", "samples", "tri_u", "tri_v", "tri_t", "tri_accept")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


MODE_PRIMARY, MODE_AO, MODE_DIRECT, MODE_PATH = 0, 1, 2, 3
ACCUM_REFERENCE, ACCUM_LINEAR = 0, 1

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.ora_hash_uint32.restype = C.c_uint32
        L.ora_hash_uint32.argtypes = [C.c_uint32]
        L.ora_tan.restype = C.c_float
        L.ora_tan.argtypes = [C.c_float]
        L.ora_pow.restype = C.c_float
        L.ora_pow.argtypes = [C.c_float, C.c_float]
        L.ora_pow_array.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.ora_render.restype = C.c_int
        L.ora_load_model.restype = C.c_int
        L.ora_tessellate.restype = C.c_int
        L.ora_max_threads.restype = C.c_int
        L.ora_local_pixel_count.restype = C.c_int
        L.ora_bvh_build.restype = C.c_int
        L.ora_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def load_model(path):
    tris = np.zeros(4096, TRIANGLE_DTYPE)
    mats = np.zeros(2048, MATERIAL_DTYPE)
    nt, nm = C.c_int(0), C.c_int(0)
    rc = lib().ora_load_model(path.encode(), _p(tris), 4096, _p(mats), 2048, C.byref(nt), C.byref(nm))
    if rc != 0:
        raise RuntimeError(f"ora_load_model({path}) -> {rc}")
    return tris[: nt.value].copy(), mats[: nm.value].copy()


def tessellate(tris, k):
    out = np.zeros(len(tris) * k * k, TRIANGLE_DTYPE)
    n = lib().ora_tessellate(_p(tris), len(tris), k, _p(out), len(out))
    assert n == len(out), n
    return out


def light_from_quad(tris, quad):
    p1, ea, eb = (C.c_float * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    lib().ora_light_from_quad(_p(tris), len(tris), quad, p1, ea, eb)
    return list(p1), list(ea), list(eb)


def rng_kat(gid, frame, n):
    st = np.zeros(n, np.uint32)
    va = np.zeros(n, np.float32)
    lib().ora_rng_kat(C.c_uint32(gid), C.c_uint32(frame), n, _p(st), _p(va))
    return st, va


def sincos(x):
    x = np.ascontiguousarray(x, np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    lib().ora_sincos_array(_p(x), x.size, _p(s), _p(c))
    return s, c


def powf(x, y):
    x = np.ascontiguousarray(x, np.float32)
    o = np.empty_like(x)
    lib().ora_pow_array(_p(x), x.size, C.c_float(y), _p(o))
    return o


def generate_ray(gi, gj, w, h, seed):
    s = C.c_uint32(seed)
    o, d = (C.c_float * 3)(), (C.c_float * 3)()
    lib().ora_generate_ray(gi, gj, w, h, C.byref(s), o, d)
    return np.array(o, np.float32), np.array(d, np.float32), s.value


def make_bvh(nodes, tri_order, quantized=True):
    """nodes: NODE_DTYPE (binary) or NODE4_DTYPE (4-wide) array; tri_order: int32 array.  Returns (Bvh, keepalive)."""
    nodes = np.ascontiguousarray(nodes)
    order = np.ascontiguousarray(tri_order, np.int32)
    width = {64: 2, 128: 4, 32: 1}[nodes.dtype.itemsize]
    b = Bvh(nodes.ctypes.data, len(nodes), order.ctypes.data, len(order), width)
    keep = [nodes, order]
    if width == 2 and quantized:  # the product traverses binary trees through their quantised encoding (ptb_bvh_nodeq)
        q, lo, step = quantize(nodes)
        b.qnodes = q.ctypes.data
        b.q_lo[:] = lo
        b.q_step[:] = step
        keep.append(q)
    return b, tuple(keep)


def quantize(nodes):
    """ora_bvh_quantize: (uint32[n, 8] quantised records, grid lo[3], grid step[3]) of a binary tree"""
    nodes = np.ascontiguousarray(nodes)
    q = np.zeros((len(nodes), 8), np.uint32)
    lo, step = (C.c_float * 3)(), (C.c_float * 3)()
    if lib().ora_bvh_quantize(_p(nodes), len(nodes), _p(q), lo, step) != 0:
        raise RuntimeError("ora_bvh_quantize failed")
    return q, [float(x) for x in lo], [float(x) for x in step]


def bvh_params(**kw):
    p = BvhParams()
    lib().ora_bvh_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build_bvh(tris, params=None, width=None):
    """The oracle's own builder (oracle_bvh.c).  width None = the scene-class rule of the product: FLAT (1) for <= 64
    triangles in <= 32 leaves, else 4-wide when the 4-wide nodes + triangles + materials of the scene fit 32 KB of
    shared memory, else binary.
    Returns {"nodes", "tri_order", "depth", "bfs_nodes", "width"}."""
    tris = np.ascontiguousarray(tris)

    def one(w):
        nodes_p, order_p = C.c_void_p(), C.c_void_p()
        nn, depth, bfs = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = lib().ora_bvh_build(_p(tris), len(tris), C.byref(params) if params is not None else None, w, C.byref(nodes_p),
                                 C.byref(nn), C.byref(order_p), C.byref(depth), C.byref(bfs))
        if rc != 0:
            raise RuntimeError(f"ora_bvh_build(width={w}) -> {rc}")
        dt = {4: NODE4_DTYPE, 2: NODE_DTYPE, 1: LEAFBOX_DTYPE}[w]
        nodes = np.frombuffer(C.string_at(nodes_p, nn.value * dt.itemsize), dt).copy()
        order = np.frombuffer(C.string_at(order_p, len(tris) * 4), np.int32).copy()
        lib().ora_free(nodes_p)
        lib().ora_free(order_p)
        return {"nodes": nodes, "tri_order": order, "depth": depth.value, "bfs_nodes": bfs.value, "width": w}

    if width is not None:
        return one(width)
    if len(tris) <= 64:
        try:
            return one(1)
        except RuntimeError:
            pass  # more than 32 leaves
    if len(tris) <= 2048:
        b4 = one(4)
        n_mats = int(tris["id"].max()) + 1
        if len(b4["nodes"]) * 128 + len(tris) * 48 + n_mats * 32 <= 32 * 1024 and b4["bfs_nodes"] == len(b4["nodes"]):
            return b4
    return one(2)


def trace(tris, o, d, tmax, bvh=None, any_hit=False):
    n = len(o)
    o = np.ascontiguousarray(o, np.float32)
    d = np.ascontiguousarray(d, np.float32)
    tmax = np.ascontiguousarray(np.broadcast_to(np.asarray(tmax, np.float32), (n,)))
    out = {
        "tri": np.empty(n, np.int32), "t": np.empty(n, np.float32), "u": np.empty(n, np.float32),
        "v": np.empty(n, np.float32), "visits": np.empty(n, np.uint32), "tests": np.empty(n, np.uint32),
    }
    lib().ora_trace(_p(tris), len(tris), C.byref(bvh) if bvh is not None else None, 1 if bvh is not None else 0,
                    1 if any_hit else 0, n, _p(o), _p(d), _p(tmax), _p(out["tri"]), _p(out["t"]), _p(out["u"]),
                    _p(out["v"]), _p(out["visits"]), _p(out["tests"]))
    return out


def default_params(width=512, height=512, **kw):
    p = Params()
    p.width, p.height = width, height
    p.first_frame, p.n_frames = 0, 1
    p.mode, p.accum, p.use_bvh = MODE_PATH, ACCUM_REFERENCE, 0
    p.max_depth, p.ao_samples, p.ao_max_dist = 16, 16, 2.0
    p.light_quad = 5
    p.shard_index, p.shard_count, p.shard_block = 0, 1, 64
    p.n_threads = 0
    for k, v in kw.items():
        if k in ("light_p1", "light_ea", "light_eb"):
            getattr(p, k)[:] = v
        else:
            setattr(p, k, v)
    return p


def render(params, tris, mats, bvh=None, fb=None, want_stats=False):
    """Returns (fb[npix_local,4] float32, stats or None, counters dict)."""
    L = lib()
    n_local = L.ora_local_pixel_count(C.byref(params))
    if fb is None:
        fb = np.zeros((n_local, 4), np.float32)
    stats = np.zeros(n_local, STATS_DTYPE) if want_stats else None
    ctr = Counters()
    rc = L.ora_render(C.byref(params), _p(tris), len(tris), _p(mats), len(mats),
                      C.byref(bvh) if bvh is not None else None, _p(fb), _p(stats), C.byref(ctr))
    if rc != 0:
        raise RuntimeError(f"ora_render -> {rc}")
    return fb, stats, ctr.as_dict()


def to_rgb8(fb):
    fb = np.ascontiguousarray(fb, np.float32)
    n = fb.shape[0]
    out = np.empty((n, 3), np.uint8)
    lib().ora_to_rgb8(_p(fb), n, _p(out))
    return out


def max_threads():
    return lib().ora_max_threads()
