"""GPU: state I/O at the seam.  sg_ball2d_state_serialize writes Ball2DState::serialize's byte stream
(ball2d/Ball2DState.cpp:259-272) from the device-resident state:
  * the tail (fixed flags, drums, planes, portals, forces) is compared byte for byte with what the reference's OWN compiled classes write
    (oracle/_ref/libref_ball2d.so: Utilities.cpp, StaticDrum.cpp, StaticPlane.cpp, PlanarPortal.cpp, Ball2DGravityForce.cpp)
  * the head (q, v, r, M, Minv) is decoded with the layout of scisim/Math/MathUtilities.h:42-60 / MathUtilities.cpp:142-154
  * run-vs-resume bit-identity, the reference's own integration-test idea (assets/*/shell_scripts/execute_serialization_test.sh):
    a sim resumed from the snapshot steps exactly like the one that wrote it."""
import ctypes as C
import os

import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref_ball2d.so")


def _head(blob, n):
    off = 0
    out = {}
    for name, cnt in (("q", 2 * n), ("v", 2 * n), ("r", n)):
        assert np.frombuffer(blob, np.int64, 1, off)[0] == cnt
        out[name] = np.frombuffer(blob, np.float64, cnt, off + 8)
        off += 8 + 8 * cnt
    assert np.frombuffer(blob, np.uint64, 1, off)[0] == n
    tail_start = off
    off += 8 + n
    for name in ("M", "Minv"):
        rows, cols, nnz = np.frombuffer(blob, np.int64, 3, off)
        assert rows == cols == nnz == 2 * n
        off += 24
        assert np.array_equal(np.frombuffer(blob, np.int32, 2 * n, off), np.arange(2 * n)); off += 8 * n
        assert np.array_equal(np.frombuffer(blob, np.int32, 2 * n + 1, off), np.arange(2 * n + 1)); off += 4 * (2 * n + 1)
        out[name] = np.frombuffer(blob, np.float64, 2 * n, off); off += 16 * n
    return out, tail_start, off


def test_snapshot_layout_and_resume(oracle, gpu_ctx):
    import scisim_b200 as sb
    s = scenes.ball2d_periodic(3000, 5, axes="x", lees_edwards=0.7, t=0.3)
    s["drum_x"], s["drum_r"] = np.array([[3.0, 4.0]]), np.array([90.0])
    s["g"] = np.array([0.3, -9.81])
    n = 3000
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"], planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    sim.updatePeriodicBoundaryConditionsStartOfStep(3, 0.1)     # t = 0.3: the Lees-Edwards portal has moved
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.SymplecticEulerMap(), s["dt"])
    q1, v1, a = sim.fetch()
    blob = sim.serializeState(which=1)
    head, tail_start, mats_end = _head(blob, n)
    assert np.array_equal(head["q"], q1) and np.array_equal(head["v"], v1) and np.array_equal(head["r"], s["r"])
    assert np.array_equal(head["M"], np.repeat(s["m"], 2)) and np.array_equal(head["Minv"], np.repeat(1.0 / s["m"], 2))
    if os.path.exists(REF):
        ref = C.CDLL(REF)
        ref.ref_ball2d_snapshot_tail.restype = C.c_uint64
        p = s["portals"]
        vp = lambda x: np.ascontiguousarray(x, dtype=np.float64).ctypes.data_as(C.c_void_p)
        out = np.zeros(1 << 20, dtype=np.uint8)
        arrs = [s["drum_x"], s["drum_r"], s["plane_x"], s["plane_n"], p["plane_a_x"], p["plane_a_n"], p["plane_b_x"], p["plane_b_n"], p["v"], p["bounds"], s["g"]]
        keep = [np.ascontiguousarray(x, dtype=np.float64) for x in arrs]
        k = [x.ctypes.data_as(C.c_void_p) for x in keep]
        nb = ref.ref_ball2d_snapshot_tail(C.c_uint32(n), C.c_uint32(1), k[0], k[1], C.c_uint32(s["plane_x"].shape[0]), k[2], k[3], C.c_uint32(p["v"].shape[0]), k[4], k[5], k[6], k[7], k[8], k[9],
                                          C.c_double(3 * 0.1), k[10], out.ctypes.data_as(C.c_void_p), C.c_uint64(out.shape[0]))
        ref_tail = out[: int(nb)].tobytes()
        fixed_len = 8 + n
        mine = blob[tail_start:tail_start + fixed_len] + blob[mats_end:]
        assert mine == ref_tail, "tail of the snapshot differs from what the reference's classes write"
    # resume: a fresh context configured from the snapshot continues exactly like the original
    ctx2 = sb.Context(0)
    sim2 = sb.Ball2DSim.deserializeState(blob, ctx2)
    blob2 = sim2.serializeState(which=0)
    assert blob2 == blob
    sim.upload(q1, v1)
    c1 = sim.step(sb.SymplecticEulerMap(), s["dt"])
    c2 = sim2.step(sb.SymplecticEulerMap(), s["dt"])
    assert c1 == c2
    qa, va, aa = sim.fetch()
    qb, vb, ab = sim2.fetch()
    assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    assert np.array_equal(aa.candidates, ab.candidates)
    for k in ("type", "i", "j", "n", "p"):
        assert np.array_equal(getattr(aa, k), getattr(ab, k)), k
    assert np.array_equal(aa.depth, ab.depth, equal_nan=True)
    ctx2.close()


def test_snapshot_is_the_references_own_byte_for_byte(oracle, gpu_ctx):
    """The whole snapshot, not only its tail: the reference's own Ball2DState::serialize (ball2d/Ball2DState.cpp compiled unchanged, the leaf serialisers of
    MathUtilities restated in its byte layout -- tests/test_reference_sim_cpu.py checks that side on the CPU) of a reference simulation holding the same
    state writes exactly the bytes sg_ball2d_state_serialize writes from the device; and the reference's Ball2DState::deserialize accepts the product's
    snapshot and writes it back unchanged."""
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    import scisim_b200 as sb
    from tests.reference_sim_binding import RefBall2DSim
    s = scenes.ball2d_periodic(2000, 6, axes="x")
    s["drum_x"], s["drum_r"] = np.array([[3.0, 4.0], [-1.0, 2.0]]), np.array([90.0, 120.0])
    s["g"] = np.array([0.3, -9.81])
    n = 2000
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"], planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    sim.upload(s["q"], s["v"])
    sim.step(sb.SymplecticEulerMap(), s["dt"])
    q1, v1, _ = sim.fetch()
    blob = sim.serializeState(which=1)
    ref = RefBall2DSim(s, s["portals"])
    ref.set_state(q1, v1)
    theirs = ref.serialize_state()
    assert len(theirs) == len(blob)
    assert theirs == blob
    again = RefBall2DSim.from_snapshot(blob, n)
    assert again.serialize_state() == blob


@pytest.mark.parametrize("scene", ["boxes_cylinders", "portals"])
def test_rb3d_snapshot_is_the_references_own_and_resumes(oracle, gpu_ctx, scene):
    """sg_rb3d_state_serialize from the device-resident rigidbody3d state (spheres and boxes; the byte layout itself is checked on the CPU against the reference's
    RigidBody3DState::serialize, tests/test_rb3d_snapshot_cpu.py): the whole snapshot equals what the reference's own RigidBody3DState writes for the same state --
    before any flow (constructor's mass-matrix layout) and after a step (updateMandMinv's) -- and a context restored from it continues exactly like the original."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rb3d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    import scisim_b200 as sb
    from tests.reference_sim_binding import RefRB3DSim
    from tests.test_rb3d_gpu import make_sim
    portals = None
    if scene == "boxes_cylinders":
        s = scenes.rb3d_random_boxes(1500, 151, spin=True, nfixed_frac=0.0, nplanes=3)
    else:
        s = scenes.rb3d_periodic_spheres(1500, 152, axes="xz")
        portals = s["portals"]
    n = s["geo_of_body"].shape[0]
    if portals is None:
        sim = make_sim(s, gpu_ctx)
    else:
        st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"],
                                 planar_portals=sb.PlanarPortal3D.from_arrays(portals))
        sim = sb.RigidBody3DSim(st, ctx=gpu_ctx)
    ref = RefRB3DSim(s, portals)
    sim.upload(s["q"], s["v"])
    assert sim.serializeState(which=0, m_updated=False) == ref.serialize_state()
    sim.step(sb.DMVMap(), s["dt"])
    q1, v1, a = sim.fetch()
    blob = sim.serializeState(which=1)
    assert blob == ref.serialize_state(q1, v1, update=True)
    # resume in a fresh context: same snapshot back, same next step
    ctx2 = sb.Context(0)
    sim2 = sb.RigidBody3DSim.deserializeState(blob, ctx2)
    assert sim2.serializeState(which=0, m_updated=True) == blob
    sim.updateMandMinv()          # the original has taken a step: every later flow reads the updated M, as the restored sim does
    sim.upload(q1, v1)
    c1 = sim.step(sb.DMVMap(), s["dt"])
    c2 = sim2.step(sb.DMVMap(), s["dt"])
    assert c1 == c2 and c1[1] > 0
    qa, va, aa = sim.fetch()
    qb, vb, ab = sim2.fetch()
    assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    for k in ("type", "i", "j", "aux", "n", "p"):
        assert np.array_equal(getattr(aa, k), getattr(ab, k)), k
    assert np.array_equal(aa.depth, ab.depth, equal_nan=True)
    ctx2.close()
