import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENE = os.path.join(ROOT, "data", "cornellbox.bin")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def built():
    """Both shared libraries exist (built by __graft_entry__.build())."""
    import oclpathtracer_b200 as pt
    from oracle import binding as ob

    if not os.path.exists(pt.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_pt.so")):
        import __graft_entry__

        __graft_entry__.build()
    return pt, ob


@pytest.fixture(scope="session")
def pt(built):
    return built[0]


@pytest.fixture(scope="session")
def ob(built):
    return built[1]


@pytest.fixture(scope="session")
def cornell(ob):
    return ob.load_model(SCENE)


@pytest.fixture(scope="session")
def cornell_bvh(ob, cornell):
    """The ORACLE's own tree for the Cornell box (oracle/oracle_bvh.c) in the form the product picks for that scene
    (FLAT: 18 leaf boxes).  Product trees are compared with it byte for byte (tests/test_host.py)."""
    tris, _ = cornell
    b = ob.build_bvh(tris)
    bvh, keep = ob.make_bvh(b["nodes"], b["tri_order"])
    return b, bvh, keep


def oracle_tree_for_mode(ob, tris, mode, cache={}):
    """The oracle's own tree in the form the product walks for `mode` on a default-built scene: FLAT scenes keep their
    4-wide tree too and use it for DIRECT (coherent rays); see ptb_scene_mode_width."""
    auto = ob.build_bvh(tris)
    width = 4 if (auto["width"] == 1 and mode == 2) else auto["width"]
    b = auto if width == auto["width"] else ob.build_bvh(tris, width=width)
    bvh, keep = ob.make_bvh(b["nodes"], b["tri_order"])
    return bvh, keep, b


def oracle_tree(ob, tris, width=None, **params):
    """(ora_bvh, keepalive, dict) built by the oracle's own builder; params = ora_bvh_params fields."""
    b = ob.build_bvh(tris, ob.bvh_params(**params) if params else None, width=width)
    bvh, keep = ob.make_bvh(b["nodes"], b["tri_order"])
    return bvh, keep, b


def assert_same_tree(nodes, order, b):
    """the product's tree (nodes, tri_order) equals the oracle's, byte for byte"""
    assert np.ascontiguousarray(nodes).view(np.uint8).tobytes() == b["nodes"].view(np.uint8).tobytes()
    assert np.array_equal(np.asarray(order), b["tri_order"])


@pytest.fixture(scope="session")
def dev(pt):
    d = pt.Device(0)
    yield d
    d.close()


def oracle_params(ob, tris, w, h, **kw):
    p1, ea, eb = ob.light_from_quad(tris, 5)
    return ob.default_params(w, h, light_p1=p1, light_ea=ea, light_eb=eb, **kw)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)
