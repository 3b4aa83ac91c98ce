"""CPU: the product's RigidBody2DState snapshot writer / parser (scisim_b200/csrc/sg_rb2d_snapshot.h -- the header sg_rb2d.cu includes, compiled here for the
host through tests/rb2d_snapshot_harness.cpp) against the reference's OWN RigidBody2DState::serialize / deserialize (rigidbody2d/RigidBody2DState.cpp compiled
unchanged into oracle/_ref; the leaf serialisers of MathUtilities restated in the reference's byte layout): the same bytes for the same state -- circles and
boxes, kinematic bodies, planes, portals (Lees-Edwards ones with their m_dx at the step's time) -- parse -> write is the identity, truncated streams are refused,
and the reference reads the product's bytes back and carries on from them.  What the GPU adds (tests/test_zzz_rb2d_state_io_gpu.py) is only where the arrays
come from."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob
from tests.reference_sim_binding import RefRB2DSim, f64, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("snap2d") / "libsnap2d_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "rb2d_snapshot_harness.cpp")], check=True)
    lib = C.CDLL(out)
    V = C.c_void_p
    lib.snap2d_serialize.restype = C.c_uint64
    lib.snap2d_serialize.argtypes = [C.c_uint32] + [V] * 5 + [C.c_uint32] + [V] * 4 + [C.c_uint32, V, V, V, C.c_uint32] + [V] * 9 + [V, C.c_uint64]
    lib.snap2d_roundtrip.restype = C.c_int
    lib.snap2d_roundtrip.argtypes = [V, C.c_uint64, V, C.c_uint64, V, V, V]
    return lib


@pytest.fixture(scope="module")
def ref_built():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rb2d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")


def _tangent(n):
    """RigidBody2DStaticPlane( x, n ): m_t = ( -n.y, n.x ) (rigidbody2d/RigidBody2DStaticPlane.cpp:10-14)."""
    n = f64(n).reshape(-1, 2)
    return np.ascontiguousarray(np.stack([-n[:, 1], n[:, 0]], axis=1))


def _portal_dx(v, bounds, t):
    """PlanarPortal::updateMovingPortals (rigidbody2d/PlanarPortal.cpp:340-345)."""
    dx = np.zeros(len(v))
    for k in range(len(v)):
        if bounds[k] != 0.0:
            rep = int(np.floor((v[k] * t + bounds[k]) / (2.0 * bounds[k])))
            dx[k] = v[k] * t - 2.0 * rep * bounds[k]
    return dx


def product_bytes(lib, s, q, v, portals=None, t=None):
    n = s["geo_of_body"].shape[0]
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
    p = portals or {"plane_a_x": np.zeros((0, 2)), "plane_a_n": np.zeros((0, 2)), "plane_b_x": np.zeros((0, 2)), "plane_b_n": np.zeros((0, 2)), "v": np.zeros(0), "bounds": np.zeros(0)}
    npo = len(p["v"])
    dx = _portal_dx(f64(p["v"]), f64(p["bounds"]), t) if t is not None else np.zeros(npo)
    gt = u32(s["geo_type"])
    k = [f64(q), f64(v), f64(s["M"]), np.ascontiguousarray(s["fixed"], dtype=np.uint8), u32(s["geo_of_body"]), gt, f64(s["geo_r"]), f64(s["geo_half"]), f64(s["g"]),
         f64(s["plane_x"]), f64(s["plane_n"]), _tangent(s["plane_n"]),
         f64(p["plane_a_x"]), f64(p["plane_a_n"]), _tangent(p["plane_a_n"]), f64(p["plane_b_x"]), f64(p["plane_b_n"]), _tangent(p["plane_b_n"]), f64(p["v"]), f64(p["bounds"]), f64(dx)]
    args = [n] + [vp(a) for a in k[:5]] + [gt.shape[0]] + [vp(a) for a in k[5:9]] + [k[9].reshape(-1, 2).shape[0], vp(k[9]), vp(k[10]), vp(k[11]), npo] + [vp(a) for a in k[12:21]]
    need = int(lib.snap2d_serialize(*args, None, 0))
    assert need > 0
    buf = np.zeros(need, dtype=np.uint8)
    assert int(lib.snap2d_serialize(*args, vp(buf), need)) == need
    return buf.tobytes()


@pytest.mark.parametrize("scene", ["circles_boxes", "kinematic_circles", "portals", "lees_edwards"])
def test_snapshot_bytes_equal_the_references(oracle, ref_built, harness, scene):
    portals, t = None, None
    if scene == "circles_boxes":
        s = scenes.rb2d_random(600, 161, nplanes=3)
    elif scene == "kinematic_circles":
        s = scenes.rb2d_random(500, 162, kinds=("circle",), nfixed_frac=0.2, nplanes=2)
        assert s["fixed"].sum() > 10
    elif scene == "portals":
        s = scenes.rb2d_periodic(400, 163, axes="xy")
        portals = s["portals"]
    else:
        s = scenes.rb2d_periodic(400, 164, axes="xy", lees_edwards=0.7)
        portals = s["portals"]
    n = s["geo_of_body"].shape[0]
    ref = RefRB2DSim(s, portals)
    q0, v0 = f64(s["q"]), f64(s["v"])
    theirs = ref.serialize_state()
    mine = product_bytes(harness, s, q0, v0, portals)
    assert len(mine) == len(theirs) and mine == theirs
    # a running simulation: the reference's own flow over two steps (portals advanced to each step's time, bodies teleported); its state, its bytes
    for it in (1, 2):
        q1, v1 = ref.flow(0, it, 1, 100)
    t = 2 * 0.01
    theirs = ref.serialize_state()
    mine = product_bytes(harness, s, q1, v1, portals, t=t if portals is not None else None)
    assert mine == theirs
    # parse -> write is the identity, and nothing is left unread
    raw = np.frombuffer(theirs, dtype=np.uint8).copy()
    out = np.zeros(raw.shape[0], dtype=np.uint8)
    nb, nn, g = C.c_uint64(0), C.c_uint32(0), np.zeros(2)
    assert harness.snap2d_roundtrip(vp(raw), raw.shape[0], vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 0
    assert int(nb.value) == raw.shape[0] and int(nn.value) == n and np.array_equal(out, raw) and np.array_equal(g, f64(s["g"]))
    # the reference reads the product's bytes and writes them back unchanged; restored, it flows and detects as the original does
    again = RefRB2DSim.from_snapshot(mine)
    assert again.n == n and again.serialize_state() == mine
    qa, va = ref.flow(1, 3, 1, 100)
    qb, vb = again.flow(1, 3, 1, 100)
    assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    a, b = ref.active_set(q1, qa), again.active_set(q1, qa)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert a["type"].shape[0] > 10


def test_truncated_and_foreign_snapshots_are_refused(oracle, ref_built, harness):
    s = scenes.rb2d_random(50, 165, nplanes=1)
    ref = RefRB2DSim(s)
    blob = np.frombuffer(ref.serialize_state(), dtype=np.uint8).copy()
    out = np.zeros(blob.shape[0], dtype=np.uint8)
    nb, nn, g = C.c_uint64(0), C.c_uint32(0), np.zeros(2)
    for cut in (3, 100, blob.shape[0] // 2, blob.shape[0] - 1):
        assert harness.snap2d_roundtrip(vp(blob), cut, vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1
    # trailing bytes: the stream holds more than a state
    longer = np.concatenate([blob, np.zeros(8, np.uint8)])
    assert harness.snap2d_roundtrip(vp(longer), longer.shape[0], vp(np.zeros(longer.shape[0], np.uint8)), longer.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1
    # a q whose length is not a multiple of three; a huge length in a short buffer
    bad = blob.copy(); bad[:8] = np.frombuffer(np.int64(3 * 50 + 1).tobytes(), np.uint8)
    assert harness.snap2d_roundtrip(vp(bad), bad.shape[0], vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1
    bad[:8] = np.frombuffer(np.int64(3 * (1 << 30)).tobytes(), np.uint8)
    assert harness.snap2d_roundtrip(vp(bad), bad.shape[0], vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1
    # an unknown geometry type
    off = 8 + 150 * 8 + 8 + 150 * 8 + 2 * (24 + 150 * 4 + 151 * 4 + 150 * 8) + 8 + 50 + 8 + 50 * 4 + 8
    assert int(np.frombuffer(blob[off:off + 4].tobytes(), np.int32)[0]) in (0, 1)
    bad = blob.copy(); bad[off:off + 4] = np.frombuffer(np.int32(7).tobytes(), np.uint8)
    assert harness.snap2d_roundtrip(vp(bad), bad.shape[0], vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1


def test_python_side_walk_of_the_layout(oracle, ref_built):
    """scisim_b200.host_api.rb2d_snapshot_counts (what RigidBody2DSim.deserializeState sizes its host-side mirror with) on the reference's own bytes."""
    from scisim_b200.host_api import rb2d_snapshot_counts
    s = scenes.rb2d_periodic(300, 166, axes="xy", lees_edwards=0.4, boxes=True)
    assert rb2d_snapshot_counts(RefRB2DSim(s, s["portals"]).serialize_state()) == (300, 2)
    s = scenes.rb2d_random(77, 167, nplanes=5)
    assert rb2d_snapshot_counts(RefRB2DSim(s).serialize_state()) == (77, 0)
