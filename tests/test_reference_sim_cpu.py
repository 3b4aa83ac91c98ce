"""CPU: the reference's OWN simulation classes, compiled unchanged into oracle/_ref (Ball2DSim.cpp + Ball2DState.cpp, RigidBody3DSim.cpp + RigidBody3DState.cpp
and everything they call; oracle/ref_shims/ref_*_sim.cpp), run here and compared with the oracle:

  * <Sim>::computeActiveSet( q0, q1, v ) as a whole -- broad phase, narrow phases, the portal branch with its teleported collisions, static geometry, in the
    order the reference emits the constraints -- against the oracle's active set: types, indices, normals, points, depths, element by element, bit for bit;
  * Ball2DSim::flow( call_back, iteration, dt, umap ) over many steps (portal motion and periodic boundaries included) against the oracle's
    flow / update_portals / enforce_portals sequence.

This pins the oracle's GLUE (emission order, dispatch, kinematic rules, the teleported-collision set), which the per-class pins of
tests/test_oracle_vs_reference.py leave restated.  The GPU parity tests compare the product with this same oracle."""
import ctypes as C

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob

from tests.reference_sim_binding import RefBall2DSim, RefRB2DSim, RefRB3DSim, f64, vp


def _same_active_set(got, want, dim):
    assert got["type"].shape[0] == want["type"].shape[0], (got["type"].shape[0], want["type"].shape[0])
    for k in ("type", "i", "j"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["n"], want["n"]) and np.array_equal(got["p"], want["p"])
    assert np.array_equal(got["depth"], want["depth"], equal_nan=True)


@pytest.mark.parametrize("seed,n,nplanes,ndrums,kind", [(61, 1500, 3, 1, 0), (62, 900, 2, 2, 1), (63, 2, 1, 0, 0), (64, 2500, 4, 0, 1)])
def test_ball2d_sim_compute_active_set(oracle, seed, n, nplanes, ndrums, kind):
    """Ball2DSim::computeActiveSet (Ball2DSim.cpp:151-173, 553-608, 730-762) on messy scenes: ball-ball (swept boxes + CCD), then drums, then planes."""
    s = scenes.ball2d_random(n, seed, nplanes=nplanes, ndrums=ndrums)
    o = ob.Ball2DOracle(s)
    ref = RefBall2DSim(s)
    q, v = s["q"].copy(), s["v"].copy()
    seen = set()
    for step in range(3):
        q1, v1 = o.flow(kind, q, v, s["dt"])
        want = o.active_set(q, q1, "grid")
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        seen |= set(int(t) for t in got["type"])
        q, v = q1, v1
    assert 0 in seen or n < 10
    assert (2 in seen or nplanes == 0) and (1 in seen or ndrums == 0)


@pytest.mark.parametrize("name,steps", [("pool_break_ten_deep", 40), ("different_friction", 12)])
def test_ball2d_sim_on_bundled_scenes(oracle, name, steps):
    """BASELINE configs[0]: the reference's own scenes through the reference's own Ball2DSim -- flow( SymplecticEulerMap ) then computeActiveSet, step after
    step -- against the oracle.  (Unconstrained: no contact response on this path, as in the GPU step.)"""
    s = scenes.ball2d_asset(name)
    den = 10080 if name == "different_friction" else 10
    assert s["dt"] == 1.0 / den
    o = ob.Ball2DOracle(s)
    ref = RefBall2DSim(s)
    q, v = s["q"].copy(), s["v"].copy()
    total = 0
    for it in range(1, steps + 1):
        q1, v1 = o.flow(0, q, v, s["dt"])
        want = o.active_set(q, q1, "grid")
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        total += got["type"].shape[0]
        rq, rv = ref.flow(0, it, 1, den)
        assert np.array_equal(rq, q1) and np.array_equal(rv, v1)
        q, v = q1, v1
    assert total > 0 or name == "different_friction"   # (6 079 balls falling side by side: no contacts in the first steps, see tests/test_config1_cpu.py)


@pytest.mark.parametrize("axes,le,oblique,seed", [("x", 0.0, False, 71), ("xy", 0.0, False, 72), ("xy", 0.8, False, 73), ("y", -1.3, True, 74), ("xy", 0.5, True, 75)])
def test_ball2d_sim_with_portals(oracle, axes, le, oblique, seed):
    """The portal branch (Ball2DSim.cpp:159-166, 327-546, 610-728): teleported boxes, the TeleportedCollision set, BallBallConstraint /
    KinematicKickBallBallConstraint from teleported centres; and Ball2DSim::flow with moving (Lees-Edwards) portals and the periodic wrap, 8 steps."""
    s = scenes.ball2d_periodic(700, seed, axes=axes, lees_edwards=le, oblique=oblique)
    o = ob.Ball2DOracle(s)
    o.set_portals(s["portals"])
    ref = RefBall2DSim(s, s["portals"])
    q, v = s["q"].copy(), s["v"].copy()
    q, v = o.enforce_portals(q, v)
    ref.set_state(q, v)
    den = 100
    assert s["dt"] == 1.0 / den
    seen = set()
    for it in range(1, 9):
        o.update_portals(it * s["dt"])
        q1, v1 = o.flow(1, q, v, s["dt"])
        want = o.active_set_portals(q, q1, "grid")
        assert want is not None
        # the reference's computeActiveSet reads the portals' current offsets: advance its own copy through its flow first
        rq, rv = ref.flow(1, it, 1, den)
        q1w, v1w = o.enforce_portals(q1, v1)
        assert np.array_equal(rq, q1w) and np.array_equal(rv, v1w)
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        seen |= set(int(t) for t in got["type"])
        q, v = q1w, v1w
    assert 0 in seen and (3 in seen or 4 in seen) and (le == 0.0 or 4 in seen) and (axes != "xy" or 3 in seen)


# ---- rigidbody3d ------------------------------------------------------------------------------------------------------------------------------------
BODY_PAIR = (10, 11, 12, 13, 19, 30)
STATIC_WITH_INDEX = (14, 15, 16, 17)


def _same_rb3d_active_set(got, want, s, q0):
    """got: from the reference's classes; want: the oracle's list (the ABI's conventions: i = first / free body, j = second body or static object)."""
    assert got["type"].shape[0] == want["type"].shape[0], (got["type"].shape[0], want["type"].shape[0])
    assert np.array_equal(got["type"], want["type"])
    t = want["type"]
    assert np.array_equal(got["i"], want["i"])
    pair = np.isin(t, BODY_PAIR)
    assert np.array_equal(got["j"][pair], want["j"][pair])
    st = np.isin(t, STATIC_WITH_INDEX)
    assert np.array_equal(got["static"][st], want["j"][st])
    assert np.array_equal(got["n"], want["n"])
    exact = np.isin(t, (10, 14, 15, 17, 19))
    assert np.array_equal(got["p"][exact], want["p"][exact])
    # body-body and plane-body classes keep lever arms ( p - x0 ) and report x0 + arm: the point handed to their constructor up to rounding
    arm = np.isin(t, (12, 13, 16))
    if arm.any():
        assert np.abs(got["p"][arm] - want["p"][arm]).max() < 1e-12
    # a free sphere against a kinematic one: the list's p is the constructor argument X (the kinematic sphere's centre at q0); the class reports x_i - r_i n
    kin = np.isin(t, (11, 30))
    if kin.any():
        n = s["geo_of_body"].shape[0]
        x0 = q0[:3 * n].reshape(n, 3)
        r = s["geo_r"][s["geo_of_body"][want["i"][kin]]]
        assert np.array_equal(got["p"][kin], x0[want["i"][kin]] - r[:, None] * want["n"][kin])
    assert np.array_equal(got["depth"], want["depth"], equal_nan=True)


def _with_cylinders(s):
    x = s["q"][:3 * s["geo_of_body"].shape[0]].reshape(-1, 3)
    ext = float(np.abs(x).max())
    s["cyl_x"] = np.array([[0.1, -0.2, 0.3], [0.0, 0.0, 0.0]])
    s["cyl_axis"] = np.array([[0.2, 3.0, -0.1], [1.0, 0.1, 0.05]])
    s["cyl_r"] = np.array([0.8 * ext, 0.95 * ext])
    return s


@pytest.mark.parametrize("scene", ["spheres_kinematic_cylinders", "boxes", "meshes_cylinders", "mixed"])
def test_rb3d_sim_compute_active_set(oracle, scene):
    """RigidBody3DSim::computeActiveSet (RigidBody3DSim.cpp:250-262, 665-962, 1057-1260, 1414-1557): every body-body narrow phase the reference supports, the
    kinematic rules, planes and cylinders, in the reference's order."""
    if scene == "spheres_kinematic_cylinders":
        s = _with_cylinders(scenes.rb3d_random_spheres(1800, 81, spin=True, nfixed_frac=0.2, nplanes=3))
        expect = {10, 11, 14, 17}
    elif scene == "boxes":
        s = scenes.rb3d_random_boxes(700, 82, nfixed_frac=0.0, nplanes=3)
        expect = {12, 15}
    elif scene == "meshes_cylinders":
        s = _with_cylinders(scenes.rb3d_random_meshes(50, 83, nfixed_frac=0.15, nplanes=2))
        expect = {12, 13, 16, 18}
    else:
        s = scenes.rb3d_mixed_segregated(120)
        expect = {10, 11, 12, 13}
    o = ob.RB3DOracle(s)
    ref = RefRB3DSim(s)
    q0 = f64(s["q"])
    q1, _ = o.flow(3, q0, s["v"], s["dt"])
    want = o.active_set(q0, q1, "grid")
    assert want["supported"]
    got = ref.active_set(q0, q1)
    _same_rb3d_active_set(got, want, s, q0)
    assert expect <= set(int(t) for t in got["type"]), set(int(t) for t in got["type"])


@pytest.mark.parametrize("axes,nfixed,tilt,seed", [("x", 0.0, False, 91), ("xz", 0.15, False, 92), ("xyz", 0.0, True, 93)])
def test_rb3d_sim_with_portals(oracle, axes, nfixed, tilt, seed):
    """The portal branch (RigidBody3DSim.cpp:965-1397): teleported boxes, the TeleportedCollision set, TeleportedSphereSphereConstraint and
    KinematicObjectSphereConstraint from teleported centres, behind the un-teleported contacts."""
    s = scenes.rb3d_periodic_spheres(900, seed, axes=axes, nfixed_frac=nfixed, tilt=tilt)
    o = ob.RB3DOracle(s)
    o.set_portals(s["portals"])
    ref = RefRB3DSim(s, s["portals"])
    q0 = o.enforce_portals(f64(s["q"]))
    q1, _ = o.flow(3, q0, s["v"], s["dt"])
    want = o.active_set_portals(q0, q1, "grid")
    assert want["supported"]
    got = ref.active_set(q0, q1)
    _same_rb3d_active_set(got, want, s, q0)
    seen = set(int(t) for t in got["type"])
    assert 10 in seen and 19 in seen and (nfixed == 0.0 or 30 in seen)


@pytest.mark.parametrize("kind", [2, 3])
def test_rb3d_sim_flow_over_steps(oracle, kind):
    """RigidBody3DSim::flow( call_back, iteration, dt, umap ) for four steps of spinning boxes with anisotropic inertia: the first step reads the mass matrix
    as RigidBody3DState's constructor filled it (world-space blocks transposed), every later one as updateMandMinv left it -- the oracle's m_updated = False /
    True, the product's SG_MAP_M_UPDATED -- and the trajectories agree bit for bit only if that is followed."""
    s = scenes.rb3d_random_boxes(300, 95, spin=True, nfixed_frac=0.0, nplanes=0)
    assert s["dt"] == 1.0 / 200
    o = ob.RB3DOracle(s)
    ref = RefRB3DSim(s)
    lib = ref.lib
    lib.ref_rb3d_sim_flow.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]
    n = ref.n
    q, v = f64(s["q"]), f64(s["v"])
    differs = False
    for it in range(1, 5):
        q1, v1 = o.flow(kind, q, v, s["dt"], m_updated=(it > 1))
        rq, rv = np.zeros(12 * n), np.zeros(6 * n)
        lib.ref_rb3d_sim_flow(ref.h, kind, it, 1, 200, vp(rq), vp(rv))
        assert np.array_equal(rq, q1) and np.array_equal(rv, v1), it
        if it > 1:
            qa, va = o.flow(kind, q, v, s["dt"], m_updated=False)
            differs = differs or not (np.array_equal(qa, q1) and np.array_equal(va, v1))
        q, v = q1, v1
    assert differs   # the two layouts are observable: a later step integrated with the constructor's matrix gives other bits


# ---- rigidbody2d ------------------------------------------------------------------------------------------------------------------------------------
def _same_rb2d_active_set(got, want):
    assert got["type"].shape[0] == want["type"].shape[0], (got["type"].shape[0], want["type"].shape[0])
    assert np.array_equal(got["type"], want["type"])
    t = want["type"]
    assert np.array_equal(got["i"], want["i"])
    pair = np.isin(t, (20, 21, 22, 25, 26))
    assert np.array_equal(got["j"][pair], want["j"][pair])
    st = np.isin(t, (23, 24))
    assert np.array_equal(got["static"][st], want["j"][st])
    assert np.array_equal(got["n"], want["n"])
    exact = np.isin(t, (20, 23, 25, 26))
    assert np.array_equal(got["p"][exact], want["p"][exact])
    arm = t == 22   # BodyBodyConstraint keeps lever arms and reports x0 + arm: the point handed to its constructor up to rounding
    if arm.any():
        assert np.abs(got["p"][arm] - want["p"][arm]).max() < 1e-12
    # ( 21: the list's p is the kinematic body's position handed to the constructor; 24: the corner's body-space arm -- the classes report world-space points )
    assert np.array_equal(got["depth"], want["depth"], equal_nan=True)


@pytest.mark.parametrize("kinds,nfixed,seed,kind", [(("circle", "box"), 0.0, 101, 0), (("circle",), 0.2, 102, 1), (("box",), 0.0, 103, 1), (("circle", "box"), 0.0, 104, 0)])
def test_rb2d_sim_compute_active_set(oracle, kinds, nfixed, seed, kind):
    """RigidBody2DSim::computeActiveSet (RigidBody2DSim.cpp:184-348, 638-714): circle-circle CCD, circle-box, box-box, kinematic circles, plane contacts."""
    s = scenes.rb2d_random(1400, seed, kinds=kinds, nfixed_frac=nfixed, nplanes=3)
    o = ob.RB2DOracle(s)
    ref = RefRB2DSim(s)
    q0 = f64(s["q"])
    q1, _ = o.flow(kind, q0, s["v"], s["dt"])
    want = o.active_set(q0, q1, "grid")
    assert want["supported"]
    got = ref.active_set(q0, q1)
    _same_rb2d_active_set(got, want)
    seen = set(int(t) for t in got["type"])
    assert ("circle" not in kinds or 20 in seen or 22 in seen) and ("box" not in kinds or 22 in seen) and (nfixed == 0.0 or 21 in seen) and (23 in seen or 24 in seen)


@pytest.mark.parametrize("kind", [0, 1])
def test_rb2d_sim_flow_under_gravity(oracle, kind):
    """RigidBody2DSim::flow( call_back, iteration, dt, umap ) with a NearEarthGravityForce that is not zero, kinematic circles included, over five steps:
    the reference's own state (its M / Minv built by generateM / generateMinv, RigidBody2DState.cpp:15-43) against the oracle, bit for bit.  (The portal
    scenes above run without gravity; with this test the stand-in's finalize() + makeCompressed() sequence is exercised with forces that matter.)"""
    s = scenes.rb2d_random(700, 121 + kind, kinds=("circle",), nfixed_frac=0.15, nplanes=2)
    assert s["g"][1] != 0.0 and s["fixed"].sum() > 10
    o = ob.RB2DOracle(s)
    ref = RefRB2DSim(s)
    q, v = f64(s["q"]), f64(s["v"])
    for it in range(1, 6):
        q1, v1 = o.flow(kind, q, v, s["dt"])
        rq, rv = ref.flow(kind, it, 1, 100)
        assert np.array_equal(rq, q1) and np.array_equal(rv, v1)
        q, v = q1, v1
    free = np.repeat(s["fixed"] == 0, 3)
    assert np.abs(v - f64(s["v"]))[free].max() > 0.4   # 5 steps of g dt


@pytest.mark.parametrize("case", [dict(n=500, seed=111, side=10.0), dict(n=500, seed=112, side=10.0, lees_edwards=0.7, oblique=True), dict(n=500, seed=113, side=7.0, axes="y", lees_edwards=-0.9),
                                  dict(n=500, seed=114, side=10.0, lees_edwards=0.5, gravity=(0.4, -9.81))])
def test_rb2d_sim_with_portals(oracle, case):
    """The portal branch (RigidBody2DSim.cpp:350-636, 832-1040) and RigidBody2DSim::flow's portal bookkeeping over 8 steps without contact response
    (the generator's scenes are weightless; the last case adds a gravity, so the reference's own M / Minv matter to its flow)."""
    case = dict(case)
    gravity = case.pop("gravity", None)
    s = scenes.rb2d_periodic(**case)
    if gravity is not None:
        s["g"] = np.array(gravity, dtype=np.float64)
    o = ob.RB2DOracle(s)
    o.set_portals(s["portals"])
    ref = RefRB2DSim(s, s["portals"])
    q, v = o.enforce_portals(f64(s["q"]), f64(s["v"]))
    ref.set_state(q, v)
    assert s["dt"] == 0.01
    seen = set()
    for it in range(1, 9):
        o.update_portals(it * s["dt"])
        q1, v1 = o.flow(0, q, v, s["dt"])
        want = o.active_set_portals(q, q1, "grid")
        assert want["supported"]
        rq, rv = ref.flow(0, it, 1, 100)
        q1w, v1w = o.enforce_portals(q1, v1)
        assert np.array_equal(rq, q1w) and np.array_equal(rv, v1w)
        got = ref.active_set(q, q1)
        _same_rb2d_active_set(got, want)
        seen |= set(int(t) for t in got["type"])
        q, v = q1w, v1w
    assert 20 in seen and (25 in seen or 26 in seen)


def test_ball2d_state_snapshot_of_the_reference(oracle):
    """Ball2DState::serialize / deserialize compiled unchanged (with the leaf (de)serialisers of MathUtilities restated in the reference's byte layout,
    oracle/ref_shims/override): the snapshot decodes with the layout the product writes (tests/test_state_io_gpu.py compares the product's bytes with
    these on the GPU), and a simulation restored from it continues exactly like the one that wrote it."""
    s = scenes.ball2d_periodic(400, 121, axes="x", lees_edwards=0.0)
    s["drum_x"], s["drum_r"] = np.array([[3.0, 4.0]]), np.array([90.0])
    s["g"] = np.array([0.3, -9.81])
    n = 400
    ref = RefBall2DSim(s, s["portals"])
    for it in range(1, 4):
        ref.flow(0, it, 1, 100)
    blob = ref.serialize_state()
    off = 0
    q = None
    for cnt in (2 * n, 2 * n, n):
        assert np.frombuffer(blob, np.int64, 1, off)[0] == cnt
        q = q if q is not None else np.frombuffer(blob, np.float64, cnt, off + 8)
        off += 8 + 8 * cnt
    assert np.frombuffer(blob, np.uint64, 1, off)[0] == n
    off += 8 + n
    for name in ("M", "Minv"):
        rows, cols, nnz = np.frombuffer(blob, np.int64, 3, off)
        assert rows == cols == nnz == 2 * n
        off += 24
        assert np.array_equal(np.frombuffer(blob, np.int32, 2 * n, off), np.arange(2 * n)); off += 8 * n
        assert np.array_equal(np.frombuffer(blob, np.int32, 2 * n + 1, off), np.arange(2 * n + 1)); off += 4 * (2 * n + 1)
        vals = np.frombuffer(blob, np.float64, 2 * n, off); off += 16 * n
        assert np.array_equal(vals, np.repeat(s["m"] if name == "M" else 1.0 / s["m"], 2))
    again = RefBall2DSim.from_snapshot(blob, n)
    assert again.serialize_state() == blob
    qa, va = ref.flow(0, 4, 1, 100)
    qb, vb = again.flow(0, 4, 1, 100)
    assert np.array_equal(qa, qb) and np.array_equal(va, vb) and not np.array_equal(qa[:2 * n], q)
    _same_active_set(again.active_set(q, qa), ref.active_set(q, qa), 2)


def test_ball2d_compute_N_and_contact_bases_of_the_reference(oracle):
    """8(f2): ImpactOperatorUtilities::computeN (ImpactOperatorUtilities.cpp:10-48, compiled unchanged) on the active set of the reference's own Ball2DSim, sized
    and called as ImpactMap::flow does (ImpactMap.cpp:106-107), and Ball2DSim::computeContactBases (Ball2DSim.cpp:188-201): the pruned, column-compressed N --
    pattern and values -- and the 2x2 bases equal the oracle's assembly (oracle/assembly2d.h), which the device-side assembly is compared with on the GPU."""
    s = scenes.ball2d_random(600, 131, nplanes=3, ndrums=1)
    o = ob.Ball2DOracle(s)
    ref = RefBall2DSim(s)
    lib = ref.lib
    lib.ref_ball2d_sim_compute_N.restype = C.c_uint64
    lib.ref_ball2d_sim_compute_N.argtypes = [C.c_void_p] * 4 + [C.c_uint64, C.c_uint64] + [C.c_void_p] * 5
    q0, v0 = f64(s["q"]), f64(s["v"])
    q1, v1 = o.flow(0, q0, v0, s["dt"])
    a = o.active_set(q0, q1, "grid")
    asm = o.assemble()
    assert asm["supported"]
    na = a["type"].shape[0]
    assert na > 100 and (a["type"] == 1).any() and (a["type"] == 2).any()
    nnz = C.c_uint64(0)
    outer, inner, values, bases = np.zeros(na + 1, np.int32), np.zeros(4 * na, np.int32), np.zeros(4 * na), np.zeros(4 * na)
    # the bases depend on v only through the tangent's sign convention of the 2-D classes; the oracle builds them from the normals: pass v0 as the map does
    nc = int(lib.ref_ball2d_sim_compute_N(ref.h, vp(q0), vp(q1), vp(v0), na, 4 * na, C.byref(nnz), vp(outer), vp(inner), vp(values), vp(bases)))
    assert nc == na and int(nnz.value) == asm["n_values"].shape[0]
    k = int(nnz.value)
    assert np.array_equal(outer, asm["n_outer"]) and np.array_equal(inner[:k], asm["n_inner"]) and np.array_equal(values[:k], asm["n_values"])
    assert np.array_equal(bases, asm["bases"])


# ---- many small scenes ------------------------------------------------------------------------------------------------------------------------------
def test_fuzz_ball2d_sim_against_oracle(oracle):
    """160 small ball2d scenes (with and without portals, both maps, crowded and sparse, a few balls exactly on top of each other or exactly touching):
    the reference's own Ball2DSim::computeActiveSet against the oracle after each of three steps."""
    rng = np.random.default_rng(2024)
    total = 0
    for case in range(160):
        n = int(rng.integers(2, 260))
        kind = int(rng.integers(0, 2))
        if case % 4 == 3:
            s = scenes.ball2d_periodic(n, 1000 + case, axes=("x", "y", "xy")[case % 3], lees_edwards=float(rng.choice([0.0, 0.6, -1.1])), oblique=bool(case % 8 == 7))
            portals = s["portals"]
        else:
            s = scenes.ball2d_random(n, 1000 + case, nplanes=int(rng.integers(0, 4)), ndrums=int(rng.integers(0, 3)), vmax=float(rng.choice([0.0, 5.0, 60.0])))
            portals = None
        q, v = s["q"].copy(), s["v"].copy()
        if portals is None and n >= 6:
            qq = q.reshape(-1, 2)
            qq[1] = qq[0]                                        # coincident centres: the normal is the zero vector left as it is
            qq[3] = qq[2] + np.array([s["r"][2] + s["r"][3], 0.0])   # exactly touching
            v.reshape(-1, 2)[:4] = 0.0
        o = ob.Ball2DOracle(s)
        ref = RefBall2DSim(dict(s, q=q, v=v), portals)
        if portals is not None:
            o.set_portals(portals)
            q, v = o.enforce_portals(q, v)
            ref.set_state(q, v)
        for it in range(1, 4):
            if portals is not None:
                o.update_portals(it * s["dt"])
            q1, v1 = o.flow(kind, q, v, s["dt"])
            want = o.active_set_portals(q, q1, "grid") if portals is not None else o.active_set(q, q1, "grid")
            if want is None:
                break                                            # a ball touching both planes of one portal: the reference exits there
            rq, rv = ref.flow(kind, it, 1, 100)
            if portals is not None:
                q1w, v1w = o.enforce_portals(q1, v1)
            else:
                q1w, v1w = q1, v1
            assert np.array_equal(rq, q1w) and np.array_equal(rv, v1w), (case, it)
            got = ref.active_set(q, q1)
            _same_active_set(got, want, 2)
            total += got["type"].shape[0]
            q, v = q1w, v1w
    assert total > 8000


def test_fuzz_rb3d_and_rb2d_sims_against_oracle(oracle):
    """72 small rigidbody3d and 72 small rigidbody2d scenes of every body mix the reference supports: the reference's own computeActiveSet against the oracle."""
    total = 0
    for case in range(72):
        pick = case % 4
        if pick == 0:
            s = scenes.rb3d_random_spheres(40 + 5 * case, 2000 + case, spin=True, nfixed_frac=0.25, nplanes=case % 3)
        elif pick == 1:
            s = scenes.rb3d_random_boxes(30 + 3 * case, 2000 + case, nfixed_frac=0.0, nplanes=case % 4)
        elif pick == 2:
            s = scenes.rb3d_random_meshes(5 + case // 8, 2000 + case, nfixed_frac=0.2, nplanes=1 + case % 2)
        else:
            ax = ("x", "xz", "xyz")[case % 3]
            s = scenes.rb3d_periodic_spheres(60 + 4 * case, 2000 + case, axes=ax, nfixed_frac=0.0, tilt=(ax == "xyz"))   # (untilted, a y portal's normal is opposite to UnitY: Eigen's SVD branch of FromTwoVectors)
        portals = s.get("portals") if pick == 3 else None
        o = ob.RB3DOracle(s)
        ref = RefRB3DSim(s, portals)
        q0 = f64(s["q"])
        if portals is not None:
            o.set_portals(portals)
            q0 = o.enforce_portals(q0)
        q1, _ = o.flow(3, q0, s["v"], s["dt"])
        want = o.active_set_portals(q0, q1, "grid") if portals is not None else o.active_set(q0, q1, "grid")
        assert want["supported"], case
        got = ref.active_set(q0, q1)
        _same_rb3d_active_set(got, want, s, q0)
        total += got["type"].shape[0]
    for case in range(72):
        kinds = (("circle", "box"), ("circle",), ("box",))[case % 3]
        if case % 4 == 3:
            s = scenes.rb2d_periodic(80 + 4 * case, 3000 + case, side=8.0 + case % 5, axes=("x", "y", "xy")[case % 3], lees_edwards=(0.0, 0.8)[case % 2])
            portals = s["portals"]
        else:
            s = scenes.rb2d_random(60 + 7 * case, 3000 + case, kinds=kinds, nfixed_frac=0.2 if kinds == ("circle",) else 0.0, nplanes=case % 4)
            portals = None
        o = ob.RB2DOracle(s)
        ref = RefRB2DSim(s, portals)
        q0, v0 = f64(s["q"]), f64(s["v"])
        if portals is not None:
            o.set_portals(portals)
            q0, v0 = o.enforce_portals(q0, v0)
            o.update_portals(s["dt"])
            ref.set_state(q0, v0)
            ref.flow(0, 1, 1, 100)       # advances the reference's portals to the step's time (its own state is not used below)
        q1, _ = o.flow(0, q0, v0, s["dt"])
        want = o.active_set_portals(q0, q1, "grid") if portals is not None else o.active_set(q0, q1, "grid")
        assert want["supported"], case
        got = ref.active_set(q0, q1)
        _same_rb2d_active_set(got, want)
        total += got["type"].shape[0]
    assert total > 9000


@pytest.mark.parametrize("name", ["gas", "lattice"])
def test_headline_scene_generators_through_the_reference_sim(oracle, name):
    """The bench's own scene generators at reduced size -- configs[2]'s polydisperse gas (Verlet, 4 walls) and configs[1]'s lattice pile (symplectic Euler,
    3 planes) -- through the reference's own Ball2DSim: the oracle that bench.py's parity_check compares the GPU path with equals it."""
    if name == "gas":
        s, kind, den = scenes.ball2d_gas(n=30000), 1, 1000
    else:
        s, kind, den = scenes.ball2d_lattice(150, 150), 0, 1000
    assert s["dt"] == 1.0 / den
    o = ob.Ball2DOracle(s)
    ref = RefBall2DSim(s)
    q, v = s["q"].copy(), s["v"].copy()
    n = s["r"].shape[0]
    for it in range(1, 3):
        q1, v1 = o.flow(kind, q, v, s["dt"])
        want = o.active_set(q, q1, "grid")
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        rq, rv = ref.flow(kind, it, 1, den)
        assert np.array_equal(rq, q1) and np.array_equal(rv, v1)
        assert got["type"].shape[0] > 0.8 * n
        q, v = q1, v1


# ---- where the reference prints and exits -----------------------------------------------------------------------------------------------------------
def _reference_exits(fn):
    """Runs fn in a forked child (the reference reports unsupported pairings with std::exit( EXIT_FAILURE )): True when the child did not return normally."""
    import os
    pid = os.fork()
    if pid == 0:
        try:
            null = os.open(os.devnull, os.O_WRONLY)
            os.dup2(null, 1); os.dup2(null, 2)
            fn()
        finally:
            os._exit(0)
    _, status = os.waitpid(pid, 0)
    return os.WIFSIGNALED(status) or os.WEXITSTATUS(status) != 0


def test_unsupported_verdicts_are_exactly_where_the_reference_exits(oracle):
    """The oracle's `supported = False` (which the product turns into SG_ERR_UNSUPPORTED, tests/test_rb3d_gpu.py, test_rb2d_gpu.py, test_zz_*_portals_gpu.py) against the
    reference itself: for every construction the GPU tests use -- a box on a sphere, free boxes inside a static cylinder, kinematic boxes in 2-D, boxes reaching a
    portal, kinematic circles / boxes in teleported collisions -- the reference's own computeActiveSet, run in a forked child, exits; for the supported controls it returns."""
    cases = []
    # rigidbody3d: a box moved onto a sphere (sphere-box: "bring box-sphere back up", RigidBody3DSim.cpp:711)
    s = scenes.rb3d_mixed_segregated(40)
    box0 = int(np.nonzero(s["geo_type"][s["geo_of_body"]] == 0)[0][0])
    s["q"][3 * box0:3 * box0 + 3] = s["q"][0:3]
    cases.append(("rb3d", s, None, f64(s["q"]), f64(s["q"])))
    cases.append(("rb3d", scenes.rb3d_mixed_segregated(40), None, None, None))                                   # control
    # rigidbody3d: free boxes inside a static cylinder (RigidBody3DSim.cpp:1549-1553)
    s = scenes.rb3d_random_boxes(50, 25, nplanes=0)
    s["cyl_x"] = np.zeros((1, 3)); s["cyl_axis"] = np.array([[0.0, 1.0, 0.0]]); s["cyl_r"] = np.array([100.0])
    cases.append(("rb3d", s, None, None, None))
    # rigidbody3d: boxes in a scene with portals
    s = scenes.rb3d_random_boxes(50, 3, nplanes=0)
    p = scenes.rb3d_periodic_spheres(4, 1, side=2.0)["portals"]
    cases.append(("rb3d", s, p, f64(s["q"]), f64(s["q"])))
    # rigidbody2d: kinematic boxes
    s = scenes.rb2d_random(300, 5, kinds=("box",), box=2.0)
    s["fixed"][:] = 0
    s["fixed"][::7] = 1
    cases.append(("rb2d", s, None, f64(s["q"]), f64(s["q"])))
    cases.append(("rb2d", scenes.rb2d_random(300, 5, kinds=("box",), box=2.0, nfixed_frac=0.0), None, None, None))  # control
    # rigidbody2d: boxes reaching the portals; kinematic circles in teleported collisions
    s = scenes.rb2d_periodic(300, 8, side=9.0, boxes=True)
    q = s["q"].reshape(-1, 3)
    q[:, :2] = np.random.default_rng(2).uniform(0.0, s["side"], size=q[:, :2].shape)
    s["q"] = q.ravel().copy()
    cases.append(("rb2d", s, s["portals"], f64(s["q"]), f64(s["q"])))
    s = scenes.rb2d_periodic(300, 9, side=9.0)
    s["fixed"][::3] = 1
    cases.append(("rb2d", s, s["portals"], f64(s["q"]), f64(s["q"])))
    cases.append(("rb2d", scenes.rb2d_periodic(300, 9, side=9.0), scenes.rb2d_periodic(300, 9, side=9.0)["portals"], None, None))   # control
    seen = {True: 0, False: 0}
    for sim, s, portals, q0, q1 in cases:
        if sim == "rb3d":
            o = ob.RB3DOracle(s)
            if portals is not None:
                o.set_portals(portals)
            if q0 is None:
                q0 = f64(s["q"])
                q1, _ = o.flow(3, q0, s["v"], s["dt"])
            want = o.active_set_portals(q0, q1, "grid") if portals is not None else o.active_set(q0, q1, "grid")
            exits = _reference_exits(lambda: RefRB3DSim(s, portals).active_set(q0, q1))
        else:
            o = ob.RB2DOracle(s)
            if portals is not None:
                o.set_portals(portals)
                o.update_portals(0.0)
            if q0 is None:
                q0 = f64(s["q"])
                q1, _ = o.flow(0, q0, s["v"], s["dt"])
            want = o.active_set_portals(q0, q1, "grid") if portals is not None else o.active_set(q0, q1, "grid")
            exits = _reference_exits(lambda: RefRB2DSim(s, portals).active_set(q0, q1))
        assert exits == (not want["supported"]), (sim, exits, want["supported"])
        seen[exits] += 1
    assert seen[True] >= 5 and seen[False] >= 3, seen
