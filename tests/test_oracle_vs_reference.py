"""CPU: the oracle's restatement against the reference's OWN source files, compiled unchanged from /root/reference against
the Eigen stand-in in oracle/eigen_standin (oracle/Makefile.ref -> oracle/_ref/*.so, built by __graft_entry__.build() in
the container that has the reference; the libraries travel with the snapshot).  Skipped where oracle/_ref is absent.

Covered: the three broad-phase grids (ball2d/SpatialGridDetector.cpp, rigidbody3d/SpatialGridDetector.cpp,
rigidbody2d/SpatialGrid.cpp), ball-ball CCD (scisim/CollisionDetection/CollisionDetectionUtilities.cpp), the 3-D box-box
routine (rigidbody3d/Constraints/BoxBoxUtilities.cpp) and the 2-D box-box / circle-box routines
(rigidbody2d/BoxBoxTools.cpp, CircleBoxTools.cpp).  Everything is compared bit for bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def _load(name):
    path = os.path.join(REFDIR, name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/%s not built (no /root/reference in this container)" % name)
    return C.CDLL(path)


def _ref_pairs(fn, n, boxes, *extra):
    fn.restype = C.c_uint64
    cnt = int(fn(C.c_uint32(n), vp(boxes), *extra, None, C.c_uint64(0)))
    out = np.zeros((max(cnt, 1), 2), dtype=np.uint32)
    fn(C.c_uint32(n), vp(boxes), *extra, vp(out), C.c_uint64(cnt))
    return out[:cnt]


def _random_boxes(rng, n, dim, side, ext):
    lo = rng.uniform(0.0, side, size=(n, dim))
    hi = lo + rng.uniform(0.05 * ext, ext, size=(n, dim))
    return np.ascontiguousarray(np.concatenate([lo, hi], axis=1))


@pytest.mark.parametrize("n,seed", [(2, 1), (300, 2), (6000, 3)])
def test_grid_2d_ball2d_and_rb2d(oracle, n, seed):
    from tests import oracle_binding as ob
    ref = _load("libref_ball2d.so")
    ref2 = _load("libref_rb2d.so")
    rng = np.random.default_rng(seed)
    boxes = _random_boxes(rng, n, 2, np.sqrt(n) * 0.6, 1.0)
    grid = _ref_pairs(ref.ref_ball2d_overlaps, n, boxes, C.c_int(0))
    brute = _ref_pairs(ref.ref_ball2d_overlaps, n, boxes, C.c_int(1))
    rb2d = _ref_pairs(ref2.ref_rb2d_overlaps, n, boxes)
    assert np.array_equal(grid, brute) and np.array_equal(grid, rb2d)
    assert np.array_equal(ob.aabb_overlaps(boxes, "grid")[0], grid)
    assert np.array_equal(ob.aabb_overlaps(boxes, "allpairs")[0], grid)
    assert n <= 2 or grid.shape[0] > 0


def test_grid_2d_on_the_reference_fixtures(oracle):
    from tests import oracle_binding as ob
    ref = _load("libref_ball2d.so")
    fx = np.load(os.path.join(ROOT, "tests", "golden", "aabb_fixtures.npz"))
    for name in ("spatial_grid_00", "spatial_grid_01", "spatial_grid_02"):
        boxes = np.ascontiguousarray(fx[name], dtype=np.float64)
        n = boxes.shape[0]
        grid = _ref_pairs(ref.ref_ball2d_overlaps, n, boxes, C.c_int(0))
        assert np.array_equal(ob.aabb_overlaps(boxes, "grid")[0], grid)


@pytest.mark.parametrize("n,seed", [(2, 4), (500, 5), (5000, 6)])
def test_grid_3d(oracle, n, seed):
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    rng = np.random.default_rng(seed)
    boxes = _random_boxes(rng, n, 3, n ** (1.0 / 3.0) * 0.7, 1.0)
    grid = _ref_pairs(ref.ref_rb3d_overlaps, n, boxes, C.c_int(0))
    brute = _ref_pairs(ref.ref_rb3d_overlaps, n, boxes, C.c_int(1))
    assert np.array_equal(grid, brute)
    assert np.array_equal(ob.aabb_overlaps(boxes, "grid")[0], grid)
    assert n <= 2 or grid.shape[0] > 0


def test_ccd_random_and_golden(oracle):
    ref = _load("libref_ball2d.so")
    ref.ref_ball2d_ccd.restype = C.c_int
    ref.ref_ball2d_ccd.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    from tests import oracle_binding as ob
    rng = np.random.default_rng(7)
    cases = []
    for _ in range(20000):
        q0a, q0b = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        mode = rng.integers(0, 4)
        va = rng.uniform(-3, 3, 2) if mode else np.zeros(2)
        vb = rng.uniform(-3, 3, 2) if mode > 1 else va.copy()
        cases.append((q0a, q0a + va, rng.uniform(0.05, 0.6), q0b, q0b + vb, rng.uniform(0.05, 0.6)))
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "ccd_cases.json"))):
        cases.append((np.array(c["q0a"], float), np.array(c["q1a"], float), float(c["ra"]), np.array(c["q0b"], float), np.array(c["q1b"], float), float(c["rb"])))
    hits = 0
    for q0a, q1a, ra, q0b, q1b, rb in cases:
        q0a, q1a, q0b, q1b = (np.ascontiguousarray(x, dtype=np.float64) for x in (q0a, q1a, q0b, q1b))
        cr, tr = np.zeros(3), np.zeros(1)
        hr = ref.ref_ball2d_ccd(vp(q0a), vp(q1a), ra, vp(q0b), vp(q1b), rb, vp(cr), vp(tr))
        co, ho, to = ob.ccd(q0a, q1a, ra, q0b, q1b, rb)
        to = [to]
        assert np.array_equal(cr, co), (cr, co)
        assert bool(hr) == ho
        if hr:
            assert tr[0] == to[0]
            hits += 1
    assert 1000 < hits < len(cases) - 1000


def test_ball2d_detection_pipeline(oracle):
    """Swept boxes -> reference grid -> reference CCD per candidate, against the oracle's computeActiveSet."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_ball2d.so")
    s = scenes.ball2d_random(4000, 12, nplanes=0, ndrums=0)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    want = o.active_set(s["q"], q1, "grid")
    nc, na = C.c_uint64(0), C.c_uint64(0)
    q0 = np.ascontiguousarray(s["q"]); r = np.ascontiguousarray(s["r"])
    act = np.zeros((want["type"].shape[0] + 8, 2), dtype=np.uint32)
    ref.ref_ball2d_detect(C.c_uint32(4000), vp(q0), vp(q1), vp(r), C.byref(nc), C.byref(na), vp(act), C.c_uint64(act.shape[0]))
    assert nc.value == want["candidates"].shape[0] and na.value == want["type"].shape[0] and na.value > 100
    assert np.array_equal(act[:na.value, 0], want["i"]) and np.array_equal(act[:na.value, 1], want["j"])


def _rot3(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_box_box_3d(oracle):
    ref = _load("libref_rb3d.so")
    ref.ref_rb3d_box_box.restype = C.c_int
    oracle.orc_box_box_3d.restype = C.c_int
    rng = np.random.default_rng(9)
    hits, multi = 0, 0
    for k in range(6000):
        axis_aligned = k % 5 == 0     # exercises the degenerate (parallel-edge) branches
        R0 = np.eye(3) if axis_aligned else _rot3(rng)
        R1 = np.eye(3) if (axis_aligned and k % 10 == 0) else _rot3(rng)
        s0, s1 = rng.uniform(0.3, 1.5, 3), rng.uniform(0.3, 1.5, 3)
        c0 = rng.uniform(-0.2, 0.2, 3)
        c1 = c0 + rng.uniform(-1.3, 1.3, 3)
        args = [np.ascontiguousarray(a, dtype=np.float64) for a in (c0, R0.ravel(), s0, c1, R1.ravel(), s1)]
        nr, pr = np.zeros(3), np.zeros(24)
        no, po = np.zeros(3), np.zeros(24)
        kr = ref.ref_rb3d_box_box(*[vp(a) for a in args], vp(nr), vp(pr))
        ko = oracle.orc_box_box_3d(*[vp(a) for a in args], vp(no), vp(po))
        assert kr == ko, (k, kr, ko)
        assert np.array_equal(nr, no), (k, nr, no)
        assert np.array_equal(pr[:3 * kr], po[:3 * ko]), k
        hits += kr > 0
        multi += kr > 1
    assert hits > 1500 and multi > 300


def test_box_box_and_circle_box_2d(oracle):
    ref = _load("libref_rb2d.so")
    ref.ref_rb2d_box_box.restype = C.c_int
    ref.ref_rb2d_box_box.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    ref.ref_rb2d_circle_box.restype = C.c_int
    ref.ref_rb2d_circle_box.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    oracle.orc_box_box_2d.restype = C.c_int
    oracle.orc_box_box_2d.argtypes = ref.ref_rb2d_box_box.argtypes
    oracle.orc_circle_box_2d.restype = C.c_int
    oracle.orc_circle_box_2d.argtypes = ref.ref_rb2d_circle_box.argtypes
    rng = np.random.default_rng(10)
    bb_hits = cb_hits = 0
    for k in range(8000):
        x0 = np.ascontiguousarray(rng.uniform(-0.2, 0.2, 2)); x1 = np.ascontiguousarray(x0 + rng.uniform(-1.2, 1.2, 2))
        r0 = np.ascontiguousarray(rng.uniform(0.2, 0.8, 2)); r1 = np.ascontiguousarray(rng.uniform(0.2, 0.8, 2))
        t0 = 0.0 if k % 7 == 0 else float(rng.uniform(-4, 4))
        t1 = 0.0 if k % 14 == 0 else float(rng.uniform(-4, 4))
        nr, pr, no, po = np.zeros(2), np.zeros(4), np.zeros(2), np.zeros(4)
        kr = ref.ref_rb2d_box_box(vp(x0), t0, vp(r0), vp(x1), t1, vp(r1), vp(nr), vp(pr))
        ko = oracle.orc_box_box_2d(vp(x0), t0, vp(r0), vp(x1), t1, vp(r1), vp(no), vp(po))
        assert kr == ko and np.array_equal(nr, no) and np.array_equal(pr[:2 * kr], po[:2 * ko]), k
        bb_hits += kr > 0
        rad = float(rng.uniform(0.1, 0.7))
        nr, pr, no, po = np.zeros(2), np.zeros(2), np.zeros(2), np.zeros(2)
        hr = ref.ref_rb2d_circle_box(vp(x0), rad, vp(x1), t1, vp(r1), vp(nr), vp(pr))
        ho = oracle.orc_circle_box_2d(vp(x0), rad, vp(x1), t1, vp(r1), vp(no), vp(po))
        assert hr == ho, k
        if hr:
            assert np.array_equal(nr, no) and np.array_equal(pr, po), k
            cb_hits += 1
    assert bb_hits > 2000 and cb_hits > 2000


def test_geometry_aabbs_3d_and_2d(oracle):
    """computeAABB of spheres / boxes (rigidbody3d/Geometry/RigidBodySphere.cpp:45-49, RigidBodyBox.cpp:45-51) and of circles /
    boxes in 2-D (rigidbody2d/CircleGeometry.cpp:32-43, BoxGeometry.cpp:32-42: swept and at one configuration)."""
    ref3 = _load("libref_rb3d.so")
    ref2 = _load("libref_rb2d.so")
    if not hasattr(ref3, "ref_rb3d_aabb") or not hasattr(ref2, "ref_rb2d_aabb"):
        pytest.skip("oracle/_ref predates the geometry shims")
    lib = oracle
    for f, lb in ((lib.orc_rb3d_aabb, None), (ref3.ref_rb3d_aabb, None)):
        f.restype = None
        f.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for f in (lib.orc_rb2d_aabb, ref2.ref_rb2d_aabb):
        f.restype = None
        f.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    rng = np.random.default_rng(41)
    from scisim_b200.scenes import _random_rotations
    Rs = _random_rotations(rng, 3000)
    Rs[:50] = np.eye(3).ravel()  # axis aligned: zeros in |R|
    for k in range(3000):
        t = int(rng.integers(0, 2))
        half = np.ascontiguousarray(rng.uniform(0.1, 2.0, size=3))
        cm = np.ascontiguousarray(rng.uniform(-50, 50, size=3))
        R = np.ascontiguousarray(Rs[k])
        a, b = np.zeros(6), np.zeros(6)
        lib.orc_rb3d_aabb(t, 0.7, vp(half), vp(cm), vp(R), vp(a))
        ref3.ref_rb3d_aabb(t, 0.7, vp(half), vp(cm), vp(R), vp(b))
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (t, k)
    for k in range(4000):
        t, swept = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        half = np.ascontiguousarray(rng.uniform(0.1, 2.0, size=2))
        q0b = np.ascontiguousarray(rng.uniform(-9, 9, size=3))
        q1b = np.ascontiguousarray(q0b + rng.uniform(-1, 1, size=3))
        if k < 40:
            q1b[2] = [0.0, np.pi / 2, np.pi, -np.pi / 2][k % 4]
        a, b = np.zeros(4), np.zeros(4)
        lib.orc_rb2d_aabb(t, 0.45, vp(half), vp(q0b), vp(q1b), swept, vp(a))
        ref2.ref_rb2d_aabb(t, 0.45, vp(half), vp(q0b), vp(q1b), swept, vp(b))
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (t, swept, k)


def _ref_mesh(ref, m):
    ref.ref_rb3d_mesh_create.restype = C.c_void_p
    ref.ref_rb3d_mesh_create.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    a = [np.ascontiguousarray(m[k], dtype=np.float64) for k in ("verts", "samples", "hull", "cell_delta", "origin", "sdf")]
    dims = np.ascontiguousarray(m["dims"], dtype=np.uint32)
    return ref.ref_rb3d_mesh_create(a[0].shape[0], vp(a[0]), a[1].shape[0], vp(a[1]), a[2].shape[0], vp(a[2]), vp(a[3]), vp(dims), vp(a[4]), vp(a[5]))


def test_mesh_sdf_narrow_phase_against_the_reference_sources(oracle):
    """a15: RigidBodyTriangleMesh::computeAABB / detectCollision (trilinear distance + gradient on the signed distance grid) and
    MeshMeshUtilities::computeActiveSet / computeMeshHalfPlaneActiveSet -- the reference's own sources (built through the mesh class's
    stream constructor) against the restatement, bit for bit: single samples, whole mesh pairs in both directions, hull-vertex sets."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    if not hasattr(ref, "ref_rb3d_mesh_mesh"):
        pytest.skip("oracle/_ref predates the mesh shims")
    s = scenes.rb3d_random_meshes(14, 31, nplanes=2)
    o = ob.RB3DOracle(s)
    lib = oracle
    lib.orc_rb3d_mesh_detect.restype = C.c_int
    lib.orc_rb3d_mesh_detect.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.orc_rb3d_body_aabb.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    ref.ref_rb3d_mesh_detect.restype = C.c_int
    ref.ref_rb3d_mesh_detect.argtypes = [C.c_void_p] * 3
    ref.ref_rb3d_mesh_aabb.argtypes = [C.c_void_p] * 4
    ref.ref_rb3d_mesh_mesh.restype = C.c_uint64
    ref.ref_rb3d_mesh_mesh.argtypes = [C.c_void_p] * 8 + [C.c_uint64]
    ref.ref_rb3d_mesh_halfplane.restype = C.c_uint64
    ref.ref_rb3d_mesh_halfplane.argtypes = [C.c_void_p] * 6 + [C.c_uint64]
    ref.ref_rb3d_mesh_destroy.argtypes = [C.c_void_p]
    meshes = [_ref_mesh(ref, m) for m in s["meshes"]]
    rng = np.random.default_rng(12)
    # ---- detectCollision: points all over (and beyond) the grid
    hits = 0
    for mi, m in enumerate(s["meshes"]):
        lo = np.asarray(m["origin"], dtype=np.float64)
        hi = lo + (np.asarray(m["dims"]) - 1) * np.asarray(m["cell_delta"])
        for k in range(6000):
            x = np.ascontiguousarray(rng.uniform(lo - 0.1, hi + 0.1))
            a, b = np.zeros(3), np.zeros(3)
            ha = lib.orc_rb3d_mesh_detect(o.h, mi, vp(x), vp(a))
            hb = ref.ref_rb3d_mesh_detect(meshes[mi], vp(x), vp(b))
            assert ha == hb and (not ha or np.array_equal(a.view(np.uint64), b.view(np.uint64))), (mi, x)
            hits += ha
    assert hits > 500
    # ---- AABBs and whole pairs at the scene's own configuration
    n = s["geo_of_body"].shape[0]
    q = s["q"]
    X = q[: 3 * n].reshape(n, 3)
    R = q[3 * n:].reshape(n, 9)
    for b in range(n):
        a, c = np.zeros(6), np.zeros(6)
        lib.orc_rb3d_body_aabb(o.h, b, vp(np.ascontiguousarray(X[b])), vp(np.ascontiguousarray(R[b])), vp(a))
        ref.ref_rb3d_mesh_aabb(meshes[int(s["geo_mesh"][s["geo_of_body"][b]])], vp(np.ascontiguousarray(X[b])), vp(np.ascontiguousarray(R[b])), vp(c))
        assert np.array_equal(a.view(np.uint64), c.view(np.uint64)), b
    act = o.active_set(q, q, "allpairs")
    assert act["supported"]
    fixed = s["fixed"].astype(bool)
    bb = np.isin(act["type"], [12, 13])
    total = 0
    for (i, j) in act["candidates"]:
        if fixed[i] and fixed[j]:
            continue
        b0, b1 = (int(j), int(i)) if fixed[i] else (int(i), int(j))   # the kinematic body goes second
        cap = 20000
        p, nn = np.zeros((cap, 3)), np.zeros((cap, 3))
        cnt = int(ref.ref_rb3d_mesh_mesh(meshes[int(s["geo_mesh"][s["geo_of_body"][b0]])], vp(np.ascontiguousarray(X[b0])), vp(np.ascontiguousarray(R[b0])),
                                         meshes[int(s["geo_mesh"][s["geo_of_body"][b1]])], vp(np.ascontiguousarray(X[b1])), vp(np.ascontiguousarray(R[b1])), vp(p), vp(nn), cap))
        sel = bb & (act["i"] == b0) & (act["j"] == b1)
        assert int(sel.sum()) == cnt, (i, j)
        assert np.array_equal(act["p"][sel].view(np.uint64), p[:cnt].view(np.uint64)) and np.array_equal(act["n"][sel].view(np.uint64), nn[:cnt].view(np.uint64)), (i, j)
        total += cnt
    assert total == int(bb.sum()) and total > 200
    # ---- plane vs mesh: hull vertices below each plane
    pl = act["type"] == 16
    assert pl.sum() > 10
    for k in range(s["plane_x"].shape[0]):
        x0 = np.ascontiguousarray(s["plane_x"][k])
        nrm = np.ascontiguousarray(s["plane_n"][k] / np.linalg.norm(s["plane_n"][k]))
        # the oracle normalises the plane as StaticPlane does; take its stored normal from a contact when there is one
        sel = pl & (act["j"] == k)
        if sel.any():
            nrm = np.ascontiguousarray(act["n"][sel][0])
        for b in range(n):
            if fixed[b]:
                continue
            out = np.zeros(4096, dtype=np.uint32)
            cnt = int(ref.ref_rb3d_mesh_halfplane(meshes[int(s["geo_mesh"][s["geo_of_body"][b]])], vp(np.ascontiguousarray(X[b])), vp(np.ascontiguousarray(R[b])), vp(x0), vp(nrm), vp(out), 4096))
            mine = act["aux"][sel & (act["i"] == b)]
            assert np.array_equal(mine, out[:cnt]), (k, b)
    for m in meshes:
        ref.ref_rb3d_mesh_destroy(m)


@pytest.mark.parametrize("kind", [0, 1], ids=["symplectic_euler", "verlet"])
def test_ball2d_maps_against_the_reference_sources(oracle, kind):
    """a1 / a2 / a6: ball2d/SymplecticEulerMap.cpp, VerletMap.cpp and Forces/Ball2DGravityForce.cpp, compiled unchanged and driven
    through a FlowableSystem with Ball2DSim's diagonal M / Minv and computeForce ( setZero + gravity ), against the restated flow."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_ball2d.so")
    if not hasattr(ref, "ref_ball2d_flow"):
        pytest.skip("oracle/_ref predates the map shim")
    ref.ref_ball2d_flow.argtypes = [C.c_int, C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint, C.c_double, C.c_void_p, C.c_void_p]
    for n, seed in ((1, 1), (37, 2), (5000, 3)):
        s = scenes.ball2d_random(n, seed)
        o = ob.Ball2DOracle(s)
        q1, v1 = o.flow(kind, s["q"], s["v"], s["dt"])
        a = [np.ascontiguousarray(s[k], dtype=np.float64) for k in ("m", "r", "g", "q", "v")]
        rq1, rv1 = np.zeros(2 * n), np.zeros(2 * n)
        ref.ref_ball2d_flow(kind, n, vp(a[0]), vp(a[1]), vp(a[2]), vp(a[3]), vp(a[4]), 3, float(s["dt"]), vp(rq1), vp(rv1))
        assert np.array_equal(q1.view(np.uint64), rq1.view(np.uint64)) and np.array_equal(v1.view(np.uint64), rv1.view(np.uint64)), (n, seed)
        assert np.any(q1 != s["q"]) and np.any(v1 != s["v"])


@pytest.mark.parametrize("kind", [0, 1], ids=["symplectic_euler", "verlet"])
def test_rb2d_maps_against_the_reference_sources(oracle, kind):
    """a3: rigidbody2d/SymplecticEulerMap.cpp, VerletMap.cpp and NearEarthGravityForce.cpp compiled unchanged ( time argument
    ( iteration - 1 ) * dt, forces zeroed on kinematically scripted bodies ) against the restated flow."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb2d.so")
    if not hasattr(ref, "ref_rb2d_flow"):
        pytest.skip("oracle/_ref predates the map shim")
    ref.ref_rb2d_flow.argtypes = [C.c_int, C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint, C.c_double, C.c_void_p, C.c_void_p]
    for n, seed, kw in ((1, 1, {}), (60, 2, {}), (3000, 3, dict(kinds=("circle",), nfixed_frac=0.3))):
        s = scenes.rb2d_random(n, seed, **kw)
        o = ob.RB2DOracle(s)
        q1, v1 = o.flow(kind, s["q"], s["v"], s["dt"])
        M, g, q0, v0 = [np.ascontiguousarray(s[k], dtype=np.float64) for k in ("M", "g", "q", "v")]
        fixed = np.ascontiguousarray(s["fixed"], dtype=np.uint8)
        rq1, rv1 = np.zeros(3 * n), np.zeros(3 * n)
        ref.ref_rb2d_flow(kind, n, vp(M), vp(fixed), vp(g), vp(q0), vp(v0), 2, float(s["dt"]), vp(rq1), vp(rv1))
        assert np.array_equal(q1.view(np.uint64), rq1.view(np.uint64)) and np.array_equal(v1.view(np.uint64), rv1.view(np.uint64)), (n, seed)
    assert fixed.sum() > 100


# ---- rigidbody3d maps: SplitHamMap.cpp / DMVMap.cpp compiled unchanged, driven through a shim FlowableSystem -----------------------------
@pytest.mark.parametrize("kind", [2, 3])
@pytest.mark.parametrize("m_updated", [False, True])
@pytest.mark.parametrize("scene_kind,n,seed", [("spheres_nospin", 500, 1), ("boxes_spin", 600, 2), ("spheres_spin_fixed", 400, 3), ("one", 1, 4)])
def test_rb3d_maps_equal_reference_sources(oracle, kind, m_updated, scene_kind, n, seed):
    """q1, v1 of the oracle's restated SplitHamMap (kind 2) / DMVMap (kind 3) == the reference's own SplitHamMap.cpp:17-182 / DMVMap.cpp:15-207,
    bit for bit: spinning boxes with anisotropic inertia, kinematically scripted bodies, both mass-matrix layouts (the constructor's transposed
    blocks for a simulation's first flow, updateMandMinv's for the later ones)."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    if scene_kind == "boxes_spin":
        s = scenes.rb3d_random_boxes(n, seed, spin=True)
    elif scene_kind == "spheres_nospin":
        s = scenes.rb3d_random_spheres(n, seed, spin=False)
    else:
        s = scenes.rb3d_random_spheres(n, seed, spin=True, nfixed_frac=0.2 if scene_kind == "spheres_spin_fixed" else 0.0)
    if scene_kind == "spheres_spin_fixed":
        # kinematically scripted bodies are asserted to be at rest by the maps
        fx = s["fixed"].astype(bool)
        v = s["v"].copy()
        v[:3 * n].reshape(n, 3)[fx] = 0.0
        v[3 * n:].reshape(n, 3)[fx] = 0.0
        s["v"] = v
    o = ob.RB3DOracle(s)
    q1, v1 = o.flow(kind, s["q"], s["v"], s["dt"], m_updated=m_updated)
    rq1, rv1 = np.zeros_like(q1), np.zeros_like(v1)
    q0, v0 = np.ascontiguousarray(s["q"]), np.ascontiguousarray(s["v"])
    m, I0 = np.ascontiguousarray(s["m"], dtype=np.float64), np.ascontiguousarray(s["I0"], dtype=np.float64)
    fixed = np.ascontiguousarray(s["fixed"], dtype=np.uint8)
    g = np.ascontiguousarray(s["g"], dtype=np.float64)
    ref.ref_rb3d_flow(C.c_int(kind), C.c_uint32(n), vp(q0), vp(v0), vp(m), vp(I0), vp(fixed), vp(g), C.c_double(s["dt"]), C.c_int(1 if m_updated else 0), vp(rq1), vp(rv1))
    assert np.array_equal(q1, rq1), np.abs(q1 - rq1).max()
    assert np.array_equal(v1, rv1), np.abs(v1 - rv1).max()
    if scene_kind == "boxes_spin":
        assert np.abs(q1[3 * n:] - q0[3 * n:]).max() > 1e-6   # the orientations did move


def test_rb3d_exponential_euler_equals_reference_source(oracle):
    """rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.cpp:13-91 compiled unchanged.  Positions and velocities (the explicit Euler parts, the
    force evaluation, Minv * F) bit for bit.  The projected orientation U V^T: the reference calls Eigen::JacobiSVD, for which the stand-in supplies
    ITS OWN algorithm (eigen-decomposition of A^T A), the oracle a one-sided Jacobi -- two independent routes to the same unique polar factor, so
    they agree to rounding, not bit for bit; that is all that is claimed for it."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    for n, seed, m_updated in ((500, 5, False), (300, 6, True), (1, 7, False)):
        s = scenes.rb3d_random_boxes(n, seed, spin=True, nfixed_frac=0.0)
        o = ob.RB3DOracle(s)
        q1, v1 = o.flow(4, s["q"], s["v"], s["dt"], m_updated=m_updated)
        rq1, rv1 = np.zeros_like(q1), np.zeros_like(v1)
        q0, v0 = np.ascontiguousarray(s["q"]), np.ascontiguousarray(s["v"])
        m, I0 = np.ascontiguousarray(s["m"], dtype=np.float64), np.ascontiguousarray(s["I0"], dtype=np.float64)
        fixed = np.ascontiguousarray(s["fixed"], dtype=np.uint8)
        g = np.ascontiguousarray(s["g"], dtype=np.float64)
        ref.ref_rb3d_flow(C.c_int(4), C.c_uint32(n), vp(q0), vp(v0), vp(m), vp(I0), vp(fixed), vp(g), C.c_double(s["dt"]), C.c_int(1 if m_updated else 0), vp(rq1), vp(rv1))
        assert np.array_equal(q1[:3 * n], rq1[:3 * n]) and np.array_equal(v1, rv1)
        assert np.abs(q1[3 * n:] - rq1[3 * n:]).max() < 4e-15, np.abs(q1[3 * n:] - rq1[3 * n:]).max()
        assert np.abs(q1[3 * n:] - q0[3 * n:]).max() > 1e-4   # the orientations did move


def test_rb3d_update_m_and_minv_equals_reference_expression(oracle):
    """The oracle's updateMandMinv blocks == R * I0.asDiagonal() * R^T evaluated by the stand-in in the layout of RigidBody3DState.cpp:428-462."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    n = 700
    s = scenes.rb3d_random_boxes(n, 9, spin=True)
    I, Ii = ob.RB3DOracle(s).update_m_and_minv(s["q"])
    rI, rIi = np.zeros(9 * n), np.zeros(9 * n)
    q = np.ascontiguousarray(s["q"])
    m, I0 = np.ascontiguousarray(s["m"], dtype=np.float64), np.ascontiguousarray(s["I0"], dtype=np.float64)
    ref.ref_rb3d_mass_blocks(C.c_uint32(n), vp(q), vp(m), vp(I0), C.c_int(1), vp(rI), vp(rIi))
    assert np.array_equal(I, rI) and np.array_equal(Ii, rIi)


def test_ball2d_constraint_classes(oracle):
    """The reference's own constraint classes (ball2d/Constraints/{BallBall,BallStaticPlane,BallStaticDrum}Constraint.cpp + scisim/Constraints/
    Constraint.cpp, compiled unchanged): isActive, constructor -> normal, getWorldSpaceContactPoint( q0 ), penetrationDepth( q1 ), computeBasis and
    evalgradg for EVERY contact of an oracle active set -- normals, points, depths, the contact bases and the pruned N of the oracle's
    assembly must equal them bit for bit (rows a10-a12 and f2 of SURVEY.md section 8)."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_ball2d.so")
    ref.ref_ball2d_constraint_probe.restype = C.c_int
    ref.ref_ball2d_constraint_probe.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint32] + [C.c_void_p] * 7
    seen = {0: 0, 1: 0, 2: 0}
    for seed, nballs in ((5, 1500), (6, 900)):
        s = scenes.ball2d_random(nballs, seed, nplanes=3, ndrums=2)
        o = ob.Ball2DOracle(s)
        q0 = np.ascontiguousarray(s["q"], dtype=np.float64)
        q1, _ = o.flow(1, q0, s["v"], s["dt"])
        a = o.active_set(q0, q1, "grid")
        asm = o.assemble()
        assert asm["supported"]
        r = np.ascontiguousarray(s["r"], dtype=np.float64)
        out, rows, vals = np.zeros(10), np.zeros(8, dtype=np.int32), np.zeros(8)
        for k in range(a["type"].shape[0]):
            t, i, j = int(a["type"][k]), int(a["i"][k]), int(a["j"][k])
            if t == 1:
                geo = np.concatenate([s["drum_x"][j], [s["drum_r"][j]]]).astype(np.float64)
                kind = 2
            elif t == 2:
                geo = np.concatenate([s["plane_x"][j], s["plane_n"][j]]).astype(np.float64)
                kind = 1
            else:
                geo, kind = np.zeros(4), 0
            geo = np.ascontiguousarray(geo)
            nt = ref.ref_ball2d_constraint_probe(kind, i, j, nballs, vp(q0), vp(q1), vp(r), vp(geo), vp(out), vp(rows), vp(vals))
            if kind != 0:
                assert out[0] == 1.0   # planes and drums enter the active set through the class's isActive at q1 (Ball2DSim.cpp:730-762)
            assert np.array_equal(out[1:3], a["n"][k]) and np.array_equal(out[3:5], a["p"][k]), (t, i, j)
            assert out[5] == a["depth"][k] or (np.isnan(out[5]) and np.isnan(a["depth"][k]))
            assert np.array_equal(out[6:10], asm["bases"][4 * k: 4 * k + 4])
            # column k of the pruned N: the class's insertions with the zeros dropped (computeN prunes), rows ascending
            keep = [(int(rows[e]), float(vals[e])) for e in range(nt) if vals[e] != 0.0]
            lo, hi = int(asm["n_outer"][k]), int(asm["n_outer"][k + 1])
            assert [(int(asm["n_inner"][e]), float(asm["n_values"][e])) for e in range(lo, hi)] == sorted(keep), (t, i, j)
            seen[kind] += 1
    assert seen[0] > 200 and seen[1] > 20 and seen[2] > 5, seen


def test_rb3d_constraint_classes(oracle):
    """rigidbody3d/Constraints/{SphereSphere,StaticPlaneSphere,StaticPlaneBox}Constraint.cpp (+ scisim/Constraints/Constraint.cpp), compiled unchanged:
    isActive at q1, the constraint built as RigidBody3DSim builds it, its normal, its world-space contact point at q0 and penetrationDepth( q1 ) for
    every sphere-sphere, plane-sphere and plane-box contact of oracle active sets, bit for bit."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    ref.ref_rb3d_constraint_probe.restype = None
    ref.ref_rb3d_constraint_probe.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    seen = {10: 0, 14: 0, 15: 0}
    for s in (scenes.rb3d_random_spheres(900, 31, spin=True, nfixed_frac=0.0, nplanes=3), scenes.rb3d_random_boxes(500, 32, nfixed_frac=0.0, nplanes=3)):
        o = ob.RB3DOracle(s)
        nb = s["geo_of_body"].shape[0]
        q0 = np.ascontiguousarray(s["q"], dtype=np.float64)
        q1, _ = o.flow(1, q0, s["v"], s["dt"])
        a = o.active_set(q0, q1, "grid")
        assert a["supported"]
        out = np.zeros(8)
        for k in range(a["type"].shape[0]):
            t, i, j, aux = int(a["type"][k]), int(a["i"][k]), int(a["j"][k]), int(a["aux"][k])
            gi = int(s["geo_of_body"][i])
            if t == 10:
                geo, kind = np.array([s["geo_r"][gi], s["geo_r"][int(s["geo_of_body"][j])], 0, 0, 0, 0, 0, 0, 0], dtype=np.float64), 0
            elif t == 14:
                geo, kind = np.concatenate([s["plane_x"][j], s["plane_n"][j], [s["geo_r"][gi], 0, 0]]).astype(np.float64), 1
            elif t == 15:
                geo, kind = np.concatenate([s["plane_x"][j], s["plane_n"][j], s["geo_half"][gi]]).astype(np.float64), 2
            else:
                continue
            geo = np.ascontiguousarray(geo)
            ref.ref_rb3d_constraint_probe(kind, i, j, aux, nb, vp(q0), vp(q1), vp(geo), vp(out))
            assert out[0] == 1.0, (t, i, j, aux)
            assert np.array_equal(out[1:4], a["n"][k]) and np.array_equal(out[4:7], a["p"][k]), (t, i, j, aux, out, a["n"][k], a["p"][k])
            assert out[7] == a["depth"][k] or (np.isnan(out[7]) and np.isnan(a["depth"][k])), (t, out[7], a["depth"][k])
            seen[t] += 1
    assert seen[10] > 100 and seen[14] > 10 and seen[15] > 10, seen


def test_rb3d_cylinder_and_kinematic_sphere_constraint_classes(oracle):
    """rigidbody3d/StaticGeometry/StaticCylinder.cpp, Constraints/{StaticCylinderSphere,StaticCylinderBody,KinematicObjectSphere}Constraint.cpp compiled
    unchanged: for every cylinder-sphere, cylinder-hull-vertex and free-sphere-vs-kinematic-sphere contact of oracle active sets -- isActive at q1
    (the sphere classes), the constraint built as RigidBody3DSim builds it (RigidBody3DSim.cpp:803-816, 1504-1557), its normal and its world-space
    contact point at q0, bit for bit."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    ref.ref_rb3d_constraint_probe.restype = None
    ref.ref_rb3d_constraint_probe.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    seen = {11: 0, 17: 0, 18: 0}
    sph = scenes.rb3d_random_spheres(1500, 51, spin=True, nfixed_frac=0.25, nplanes=0)
    mesh = scenes.rb3d_random_meshes(60, 52, nfixed_frac=0.0, nplanes=0)
    for s in (sph, mesh):
        x = s["q"][:3 * s["geo_of_body"].shape[0]].reshape(-1, 3)
        ext = float(np.abs(x).max())
        s["cyl_x"] = np.array([[0.1, -0.2, 0.3], [0.0, 0.0, 0.0]])
        s["cyl_axis"] = np.array([[0.2, 3.0, -0.1], [1.0, 0.1, 0.05]])
        s["cyl_r"] = np.array([0.8 * ext, 0.95 * ext])
        o = ob.RB3DOracle(s)
        nb = s["geo_of_body"].shape[0]
        q0 = np.ascontiguousarray(s["q"], dtype=np.float64)
        q1, _ = o.flow(3, q0, s["v"], s["dt"])
        a = o.active_set(q0, q1, "grid")
        assert a["supported"]
        out = np.zeros(8)
        for k in range(a["type"].shape[0]):
            t, i, j, aux = int(a["type"][k]), int(a["i"][k]), int(a["j"][k]), int(a["aux"][k])
            gi = int(s["geo_of_body"][i])
            if t == 17:
                geo, kind = np.concatenate([s["cyl_x"][j], s["cyl_axis"][j], [s["cyl_r"][j], s["geo_r"][gi], 0, 0]]), 3
            elif t == 18:
                geo, kind = np.concatenate([s["cyl_x"][j], s["cyl_axis"][j], [s["cyl_r"][j]], a["p"][k]]), 4
            elif t == 11:
                geo, kind = np.array([s["geo_r"][gi], s["geo_r"][int(s["geo_of_body"][j])], 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.float64), 5
            else:
                continue
            geo = np.ascontiguousarray(geo, dtype=np.float64)
            ref.ref_rb3d_constraint_probe(kind, i, j, aux, nb, vp(q0), vp(q1), vp(geo), vp(out))
            assert out[0] == 1.0, (t, i, j, aux)
            assert np.array_equal(out[1:4], a["n"][k]), (t, i, j, out[1:4], a["n"][k])
            if t == 17:
                assert np.array_equal(out[4:7], a["p"][k]), (t, i, j, out[4:7], a["p"][k])
            elif t == 18:   # StaticCylinderBodyConstraint reports no contact point of its own (the base class exits); its normal is column 0 of computeContactBasis
                pass
            else:           # the list's p is the constructor argument X (the kinematic sphere's centre at q0); the class reports x_i - r_i n
                assert np.array_equal(a["p"][k], q0[3 * j: 3 * j + 3]) and np.array_equal(out[4:7], q0[3 * i: 3 * i + 3] - s["geo_r"][gi] * a["n"][k])
            seen[t] += 1
    assert seen[11] > 50 and seen[17] > 20 and seen[18] > 20, seen


def test_rb3d_state_mass_matrices_equal_reference_source(oracle):
    """rigidbody3d/RigidBody3DState.cpp compiled unchanged: the world-space blocks of M and Minv as setState builds them (formWorldSpaceMassMatrix /
    formWorldSpaceInverseMassMatrix, :140-240 -- stored TRANSPOSED, M( r, c ) = I( c, r )) and as updateMandMinv (:428-462) leaves them after the
    orientations moved, against the oracle's two layouts ( m_updated = False / True of its maps; update_m_and_minv ), bit for bit.  This is the pin for
    the 'two mass matrices' behaviour of DESIGN section 2 and for SG_MAP_M_UPDATED."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    ref.ref_rb3d_state_mass_matrices.restype = None
    ref.ref_rb3d_state_mass_matrices.argtypes = [C.c_uint32] + [C.c_void_p] * 8
    n = 600
    s = scenes.rb3d_random_boxes(n, 91, spin=True, nfixed_frac=0.0, nplanes=0)
    o = ob.RB3DOracle(s)
    q0 = np.ascontiguousarray(s["q"], dtype=np.float64)
    q1, _ = o.flow(2, q0, s["v"], s["dt"])      # orientations orthonormal only up to rounding, as in a running simulation
    q1 = np.ascontiguousarray(q1)
    m, I0 = np.ascontiguousarray(s["m"], dtype=np.float64), np.ascontiguousarray(s["I0"], dtype=np.float64)
    Mc, Mic, Mu, Miu = (np.zeros(9 * n) for _ in range(4))
    ref.ref_rb3d_state_mass_matrices(n, vp(q0), vp(q1), vp(m), vp(I0), vp(Mc), vp(Mic), vp(Mu), vp(Miu))
    # after the update: the oracle's restatement of updateMandMinv at q1
    I, Ii = o.update_m_and_minv(q1)
    assert np.array_equal(Mu, I) and np.array_equal(Miu, Ii)
    # as constructed: the same products at q0, each 3x3 block transposed
    I, Ii = o.update_m_and_minv(q0)
    T = lambda a: a.reshape(n, 3, 3).transpose(0, 2, 1).reshape(-1)
    assert np.array_equal(Mc, T(I)) and np.array_equal(Mic, T(Ii))
    assert not np.array_equal(Mc, I)            # I = R I0 R^T is symmetric only up to rounding: the two layouts do differ
    # and the shim system the map tests drive (which restates the fill) holds exactly these values
    rI, rIi = np.zeros(9 * n), np.zeros(9 * n)
    ref.ref_rb3d_mass_blocks(C.c_uint32(n), vp(q0), vp(m), vp(I0), C.c_int(0), vp(rI), vp(rIi))
    assert np.array_equal(rI, Mc) and np.array_equal(rIi, Mic)
    ref.ref_rb3d_mass_blocks(C.c_uint32(n), vp(q1), vp(m), vp(I0), C.c_int(1), vp(rI), vp(rIi))
    assert np.array_equal(rI, Mu) and np.array_equal(rIi, Miu)


def test_rb3d_update_m_and_minv_expressions(oracle):
    """RigidBody3DState::updateMandMinv (RigidBody3DState.cpp:428-462) as an EXPRESSION pin (the file itself is driven by
    test_rb3d_state_mass_matrices_equal_reference_source): its two assignments ( R * I0.asDiagonal() * R.transpose() into column-major maps ), typed as in the reference and evaluated by the stand-in, against
    the oracle's restatement, for spinning boxes at random orientations."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb3d.so")
    ref.ref_rb3d_update_inertia_expr.restype = None
    ref.ref_rb3d_update_inertia_expr.argtypes = [C.c_void_p] * 5
    s = scenes.rb3d_random_boxes(400, 77, nfixed_frac=0.0, nplanes=0)
    o = ob.RB3DOracle(s)
    n = s["geo_of_body"].shape[0]
    q1, _ = o.flow(1, s["q"], s["v"], s["dt"])   # rotations that are orthonormal only up to rounding, as in a running simulation
    I, Ii = o.update_m_and_minv(q1)
    I0 = np.ascontiguousarray(s["I0"], dtype=np.float64).reshape(n, 3)
    a, b = np.zeros(9), np.zeros(9)
    for k in range(n):
        R = np.ascontiguousarray(q1[3 * n + 9 * k: 3 * n + 9 * k + 9])
        i0 = np.ascontiguousarray(I0[k]); ii0 = np.ascontiguousarray(1.0 / I0[k])
        ref.ref_rb3d_update_inertia_expr(vp(R), vp(i0), vp(ii0), vp(a), vp(b))
        assert np.array_equal(a, I[9 * k: 9 * k + 9]) and np.array_equal(b, Ii[9 * k: 9 * k + 9]), k


def test_rb2d_constraint_classes(oracle):
    """rigidbody2d/{CircleCircle,StaticPlaneCircle,StaticPlaneBody,BodyBody}Constraint.cpp (+ scisim/Constraints/Constraint.cpp), compiled unchanged:
    the constraint built as RigidBody2DSim builds it -- normal, world-space contact point at q0, penetrationDepth( q1 ) (NaN where the class has no
    override) -- for every contact of an oracle active set; StaticPlaneCircleConstraint::isActive at q1 for the plane-circle contacts."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _load("libref_rb2d.so")
    ref.ref_rb2d_constraint_probe.restype = None
    ref.ref_rb2d_constraint_probe.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    seen = {20: 0, 22: 0, 23: 0, 24: 0}
    for seed in (41, 42):
        s = scenes.rb2d_random(1200, seed, nfixed_frac=0.0, nplanes=3)
        o = ob.RB2DOracle(s)
        nb = s["geo_of_body"].shape[0]
        q0 = np.ascontiguousarray(s["q"], dtype=np.float64)
        q1, _ = o.flow(0, q0, s["v"], s["dt"])
        a = o.active_set(q0, q1, "grid")
        assert a["supported"]
        out = np.zeros(6)
        for k in range(a["type"].shape[0]):
            t, i, j = int(a["type"][k]), int(a["i"][k]), int(a["j"][k])
            gi = int(s["geo_of_body"][i])
            if t == 20:
                geo, kind = np.array([s["geo_r"][gi], s["geo_r"][int(s["geo_of_body"][j])], 0, 0, 0, 0], dtype=np.float64), 0
            elif t == 23:
                geo, kind = np.concatenate([s["plane_x"][j], s["plane_n"][j], [s["geo_r"][gi], 0]]).astype(np.float64), 1
            elif t == 24:
                geo, kind = np.concatenate([s["plane_x"][j], s["plane_n"][j], a["p"][k]]).astype(np.float64), 2   # p carries the corner's body-space arm
            elif t == 22:
                geo, kind = np.concatenate([a["p"][k], a["n"][k], [0, 0]]).astype(np.float64), 3
            else:
                continue
            geo = np.ascontiguousarray(geo)
            ref.ref_rb2d_constraint_probe(kind, i, j, nb, vp(q0), vp(q1), vp(geo), vp(out))
            assert np.array_equal(out[1:3], a["n"][k]), (t, i, j, out, a["n"][k])
            if t == 23:
                assert out[0] == 1.0
            if t == 22:
                # the class keeps body-space arms and re-derives the point from q0: equal to the constructor's p up to rounding
                assert np.allclose(out[3:5], a["p"][k], rtol=0, atol=1e-12 * (1.0 + np.abs(a["p"][k]).max()))
            elif t != 24:
                assert np.array_equal(out[3:5], a["p"][k]), (t, i, j, out, a["p"][k])
            else:
                # the class's world-space point of the corner: x0 + R( theta0 ) * arm
                th = q0[3 * i + 2]
                c_, s_ = np.cos(th), np.sin(th)
                arm = a["p"][k]
                want = np.array([q0[3 * i] + (c_ * arm[0] - s_ * arm[1]), q0[3 * i + 1] + (s_ * arm[0] + c_ * arm[1])])
                assert np.allclose(out[3:5], want, rtol=0, atol=1e-12 * (1.0 + np.abs(want).max()))
            assert out[5] == a["depth"][k] or (np.isnan(out[5]) and np.isnan(a["depth"][k])), (t, out[5], a["depth"][k])
            seen[t] += 1
    assert seen[20] > 50 and seen[23] > 5 and seen[22] + seen[24] > 10, seen
