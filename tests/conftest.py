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
def cornell_bvh(pt, ob, cornell):
    tris, _ = cornell
    b = pt.build_bvh_host(tris)
    bvh, keep = ob.make_bvh(b["nodes"], b["tri_order"])
    return b, bvh, keep


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
