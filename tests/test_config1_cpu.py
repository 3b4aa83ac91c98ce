"""CPU: BASELINE configs[0] -- the reference's own bundled ball2d scenes (tests/golden/ball2d_assets.npz, parsed from the
reference's XML by tests/golden/make_ball2d_assets.py) -- through the oracle, step after step, against the reference's
compiled sources (oracle/_ref/libref_ball2d.so: SymplecticEulerMap.cpp, SpatialGridDetector.cpp,
CollisionDetectionUtilities.cpp): the same q1, v1, the same candidate count and the same active pairs in the same order."""
import ctypes as C
import os

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref_ball2d.so")


def test_fixture_matches_the_xml_when_the_reference_is_mounted():
    xml = "/root/reference/assets/ball2d/tests_python_serialization/different_friction.xml"
    if not os.path.exists(xml):
        pytest.skip("reference tree not mounted")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_ball2d_assets.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    s = mk.parse(xml)
    t = scenes.ball2d_asset("different_friction")
    assert s["r"].shape[0] == 6079 and np.array_equal(s["q"], t["q"]) and np.array_equal(s["r"], t["r"]) and np.array_equal(s["m"], t["m"])
    assert float(s["dt"]) == 1.0 / 10080.0 == t["dt"]


@pytest.mark.parametrize("name,steps", [("pool_break_ten_deep", 60), ("different_friction", 12)])
def test_oracle_equals_reference_sources_on_bundled_scene(oracle, name, steps):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    ref = C.CDLL(REF)
    s = scenes.ball2d_asset(name)
    n = s["r"].shape[0]
    o = ob.Ball2DOracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    r = np.ascontiguousarray(s["r"])
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    seen = 0
    for it in range(steps):
        q1, v1 = o.flow(0, q, v, s["dt"])
        a = o.active_set(q, q1, "grid")
        nc, na = C.c_uint64(0), C.c_uint64(0)
        cap = 8 * n + 64
        ij = np.zeros((cap, 2), dtype=np.uint32)
        ref.ref_ball2d_detect(C.c_uint32(n), vp(q), vp(q1), vp(r), C.byref(nc), C.byref(na), vp(ij), C.c_uint64(cap))
        bb = a["type"] == 0
        assert int(nc.value) == a["candidates"].shape[0]
        assert int(na.value) == int(bb.sum())
        assert np.array_equal(ij[: int(na.value)], np.stack([a["i"][bb], a["j"][bb]], axis=1))
        seen += int(na.value)
        q, v = q1, v1
    # (the 6 079 balls of different_friction fall side by side -- uniform gravity, no contact response on this path -- so that scene
    # only ever produces candidates and plane contacts; the cue ball of the pool break runs through the rack)
    assert seen > 0 or name == "different_friction"
