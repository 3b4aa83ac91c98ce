"""CPU: ball2d portals -- the oracle restatement (oracle/ball2d_portals.h), the reference's own
PlanarPortal.cpp / StaticPlane.cpp compiled unchanged (oracle/_ref, skipped where absent) and the product's portal
arithmetic (scisim_b200/csrc/sg_portal2d.h, the header the kernels include, compiled here as plain C++).

  * portal primitives: oracle == reference == product header, bit for bit, plain and Lees-Edwards, oblique planes
  * the product's teleported-collision sort (bitonic network + first-of-run) == std::set insertion semantics
  * the oracle's active set on a periodic box == brute-force minimum-image contacts (exact arithmetic on a dyadic grid)
  * enforcePeriodicBoundaryConditions puts every ball back inside and applies the Lees-Edwards velocity
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
vp = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def pm(tmp_path_factory):
    """The product header compiled for the host (no FMA contraction, as the library's -fmad=false)."""
    out = str(tmp_path_factory.mktemp("pm") / "libportal_math.so")
    src = os.path.join(ROOT, "tests", "portal_math_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.pm_probe.restype = C.c_uint32
    lib.pm_probe.argtypes = [C.c_uint32, C.c_void_p, C.c_double, C.c_void_p]
    lib.pm_update_portals.argtypes = [C.c_double, C.c_void_p]
    lib.pm_set_portals.argtypes = [C.c_uint32] + [C.c_void_p] * 6
    lib.pm_enforce.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    lib.pm_bitonic_sort.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    lib.pm_tele_happens.restype = C.c_int
    lib.pm_tele_happens.argtypes = [C.c_uint32] * 4 + [C.c_void_p] * 3
    return lib


def _pm_set(pm, portals):
    a = [np.ascontiguousarray(portals[k], dtype=np.float64) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n", "v", "bounds")]
    pm.pm_set_portals(a[4].shape[0], *[vp(x) for x in a])
    return a


def _oracle(scene):
    o = ob.Ball2DOracle(scene)
    o.set_portals(scene["portals"])
    return o


def _probe_points(scene, rng, count):
    side = scene["side"]
    pts = rng.uniform(-0.6 * side, 1.6 * side, size=(count, 2))
    # many points right at the planes, where the <= / < decisions are made
    edge = rng.integers(0, 4, size=count // 2)
    pts[: count // 2, 0] = np.where(edge == 0, 0.0, np.where(edge == 1, side, pts[: count // 2, 0]))
    pts[: count // 2, 1] = np.where(edge == 2, 0.0, np.where(edge == 3, side, pts[: count // 2, 1]))
    pts[: count // 4] += rng.uniform(-0.4, 0.4, size=(count // 4, 2))
    return pts


CASES = [dict(axes="xy", lees_edwards=0.0, oblique=False), dict(axes="xy", lees_edwards=0.75, oblique=False, t=3.7),
         dict(axes="x", lees_edwards=-1.25, oblique=True, t=11.3), dict(axes="y", lees_edwards=0.0, oblique=True)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-le%g-%s" % (c["axes"], c["lees_edwards"], "obl" if c["oblique"] else "axis"))
def test_portal_primitives_oracle_reference_product(case, pm, oracle):
    scene = scenes.ball2d_periodic(64, 5, **case)
    o = _oracle(scene)
    _pm_set(pm, scene["portals"])
    t = scene["t"]
    dx_o = o.update_portals(t)
    dx_p = np.zeros_like(dx_o)
    pm.pm_update_portals(t, vp(dx_p))
    assert np.array_equal(dx_o, dx_p)
    ref = None
    path = os.path.join(REFDIR, "libref_ball2d.so")
    if os.path.exists(path):
        ref = C.CDLL(path)
        if not hasattr(ref, "ref_portal_probe"):
            ref = None
    handles = []
    if ref is not None:
        ref.ref_portal_create.restype = C.c_void_p
        ref.ref_portal_create.argtypes = [C.c_void_p] * 4 + [C.c_double, C.c_double]
        ref.ref_portal_update.argtypes = [C.c_void_p, C.c_double]
        ref.ref_portal_probe.restype = C.c_uint32
        ref.ref_portal_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        ref.ref_portal_destroy.argtypes = [C.c_void_p]
        P = scene["portals"]
        for p in range(P["v"].shape[0]):
            h = ref.ref_portal_create(vp(P["plane_a_x"][p]), vp(P["plane_a_n"][p]), vp(P["plane_b_x"][p]), vp(P["plane_b_n"][p]), float(P["v"][p]), float(P["bounds"][p]))
            ref.ref_portal_update(h, t)
            handles.append(h)
    rng = np.random.default_rng(17)
    pts = _probe_points(scene, rng, 4000)
    radii = rng.uniform(0.05, 0.4, size=pts.shape[0])
    seen = set()
    for p in range(scene["portals"]["v"].shape[0]):
        for x, r in zip(pts, radii):
            x = np.ascontiguousarray(x)
            fo, oo = o.portal_probe(p, x, r)
            op = np.zeros(12)
            fp = pm.pm_probe(p, vp(x), float(r), vp(op))
            assert fo == fp and np.array_equal(oo.view(np.uint64), op.view(np.uint64)), (p, x, r)
            if ref is not None:
                orf = np.zeros(12)
                fr = ref.ref_portal_probe(handles[p], vp(x), float(r), vp(orf))
                assert fr == fo and np.array_equal(orf.view(np.uint64), oo.view(np.uint64)), (p, x, r)
            seen.add(fo)
    # the sample exercises every decision: no touch, plane A, plane B, inside, outside
    assert {0, 1, 3}.issubset({f & 3 for f in seen}) and any(f & 8 for f in seen) and any(not (f & 8) for f in seen)
    for h in handles:
        ref.ref_portal_destroy(h)
    if ref is None and os.path.isdir("/root/reference"):
        pytest.fail("oracle/_ref/libref_ball2d.so lacks the portal shim: run make -C oracle -f Makefile.ref")


def test_reference_library_has_portals_when_the_reference_is_here():
    if not os.path.isdir("/root/reference"):
        pytest.skip("no /root/reference in this container")
    lib = C.CDLL(os.path.join(REFDIR, "libref_ball2d.so"))
    assert hasattr(lib, "ref_portal_probe")


def test_bitonic_sort_is_set_insertion(pm):
    rng = np.random.default_rng(3)
    for nraw in (1, 2, 3, 7, 8, 33, 500, 1025):
        b0 = rng.integers(0, 12, size=nraw).astype(np.uint64)
        b1 = b0 + rng.integers(1, 6, size=nraw).astype(np.uint64)
        keys = (b0 << np.uint64(32)) | b1
        m = 1
        while m < nraw:
            m <<= 1
        k = np.full(m, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
        i = np.full(m, 0xFFFFFFFF, dtype=np.uint32)
        k[:nraw] = keys
        i[:nraw] = np.arange(nraw, dtype=np.uint32)
        pm.pm_bitonic_sort(m, vp(k), vp(i))
        order = np.lexsort((np.arange(nraw), keys))
        assert np.array_equal(k[:nraw], keys[order]) and np.array_equal(i[:nraw], order.astype(np.uint32))
        first = np.ones(nraw, dtype=bool)
        first[1:] = k[1:nraw] != k[: nraw - 1]
        # std::set semantics: one entry per key, the earliest inserted one
        kept = {}
        for pos, key in enumerate(keys):
            kept.setdefault(int(key), pos)
        assert [int(x) for x in i[:nraw][first]] == [kept[key] for key in sorted(kept)]


def _min_image_pairs(q, r, side, axes):
    n = r.shape[0]
    d = q[:, None, :] - q[None, :, :]
    for k, ax in enumerate("xy"):
        if ax in axes:
            d[..., k] -= side * np.round(d[..., k] / side)
    hit = (d ** 2).sum(-1) <= (r[:, None] + r[None, :]) ** 2
    i, j = np.nonzero(np.triu(hit, 1))
    return set(zip(i.tolist(), j.tolist()))


@pytest.mark.parametrize("axes", ["x", "y"])
def test_oracle_active_set_equals_minimum_image(axes, oracle):
    # dyadic coordinates and a power-of-two box: every teleport is exact, so the comparison has no rounding slack
    n, side = 700, 16.0
    rng = np.random.default_rng(9)
    scene = scenes.ball2d_periodic(n, 2, side=side, axes=axes)
    q = np.round(rng.uniform(0.0, side, size=(n, 2)) * 1024.0) / 1024.0
    q = np.minimum(q, side - 1.0 / 1024.0)
    r = np.round(rng.uniform(0.1, 0.45, size=n) * 1024.0) / 1024.0
    # keep balls off the walls of the non-periodic direction (the walls are static planes, not ball-ball business)
    scene["r"], scene["q"] = r, q.ravel().copy()
    scene["portals"]["plane_a_n"] = np.sign(scene["portals"]["plane_a_n"])
    scene["portals"]["plane_b_n"] = np.sign(scene["portals"]["plane_b_n"])
    o = _oracle(scene)
    o.update_portals(0.0)
    res = o.active_set_portals(scene["q"], scene["q"])
    assert res is not None
    bb = res["type"] != 2
    got = set(zip(res["i"][bb].tolist(), res["j"][bb].tolist()))
    assert len(got) == int(bb.sum()), "a body pair appears twice"
    want = _min_image_pairs(q, r, side, axes)
    assert got == want
    tel = np.isin(res["type"], [3, 4])
    assert tel.sum() > 3 and res["n_regular"] > 50
    # order: regular ascending, then teleported ascending, then planes
    ij = np.stack([res["i"], res["j"]], axis=1).astype(np.int64)
    nr, nt = res["n_regular"], int(tel.sum())
    for blk in (ij[:nr], ij[nr:nr + nt]):
        key = blk[:, 0] * n + blk[:, 1]
        assert np.all(np.diff(key) > 0)
    assert np.all(res["type"][:nr] == 0) and np.all(tel[nr:nr + nt]) and np.all(res["type"][nr + nt:] == 2)
    # grid == all pairs for the extended box set too
    res2 = o.active_set_portals(scene["q"], scene["q"], method="allpairs")
    assert np.array_equal(res["candidates"], res2["candidates"]) and np.array_equal(res["i"], res2["i"])


def test_both_planes_touched_is_reported(oracle):
    scene = scenes.ball2d_periodic(4, 1, side=1.0, rmin=0.6, rmax=0.7, axes="x")
    o = _oracle(scene)
    o.update_portals(0.0)
    assert o.active_set_portals(scene["q"], scene["q"]) is None


@pytest.mark.parametrize("le", [0.0, 1.5])
def test_enforce_periodic_boundary_conditions(le, pm, oracle):
    scene = scenes.ball2d_periodic(3000, 4, lees_edwards=le, t=2.3, oblique=True)
    o = _oracle(scene)
    _pm_set(pm, scene["portals"])
    o.update_portals(scene["t"])
    pm.pm_update_portals(scene["t"], None)
    rng = np.random.default_rng(8)
    q = scene["q"] + rng.uniform(-0.45, 0.45, size=scene["q"].shape) * scene["side"]
    v = scene["v"].copy()
    q1, v1 = o.enforce_portals(q, v)
    qp, vpm = q.copy(), v.copy()
    pm.pm_enforce(3000, vp(qp), vp(vpm))
    assert np.array_equal(q1, qp) and np.array_equal(v1, vpm)
    moved = np.any(q1.reshape(-1, 2) != q.reshape(-1, 2), axis=1)
    assert moved.sum() > 500
    # afterwards no ball is inside any portal
    for p in range(scene["portals"]["v"].shape[0]):
        for x in q1.reshape(-1, 2)[::7]:
            flags, _ = o.portal_probe(p, np.ascontiguousarray(x), 0.1)
            assert not (flags & 8)
    changed = np.any(v1.reshape(-1, 2) != v.reshape(-1, 2), axis=1)
    assert (changed.sum() > 100) if le != 0.0 else (changed.sum() == 0)
    assert np.all(~changed | moved)


def test_product_teleported_collision_matches_oracle(pm, oracle):
    """sg_tele_collision / sg_tele_center / sg_ball_ball_active on the oracle's own teleported candidates."""
    scene = scenes.ball2d_periodic(900, 12, side=12.0, lees_edwards=0.5, t=1.7)
    o = _oracle(scene)
    _pm_set(pm, scene["portals"])
    o.update_portals(scene["t"])
    pm.pm_update_portals(scene["t"], None)
    q0 = scene["q"]
    q1 = q0 + scene["dt"] * scene["v"]
    res = o.active_set_portals(q0, q1)
    n = scene["r"].shape[0]
    r = np.ascontiguousarray(scene["r"])
    q1c = np.ascontiguousarray(q1)
    kept = {}
    for a, b in res["candidates"]:
        ft, st = a >= n, b >= n
        if not (ft or st):
            continue
        b0, p0 = (int(res["box_body"][a - n]), int(res["box_portal"][a - n])) if ft else (int(a), 0xFFFFFFFF)
        b1, p1 = (int(res["box_body"][b - n]), int(res["box_portal"][b - n])) if st else (int(b), 0xFFFFFFFF)
        if ft and st:
            d = q1c.reshape(-1, 2)[b0] - q1c.reshape(-1, 2)[b1]
            if d[0] * d[0] + d[1] * d[1] <= (r[b0] + r[b1]) * (r[b0] + r[b1]):
                continue
        ordered = np.zeros(4, dtype=np.uint32)
        if pm.pm_tele_happens(b0, b1, p0, p1, vp(q1c), vp(r), vp(ordered)):
            kept.setdefault((int(ordered[0]), int(ordered[1])), (int(ordered[2]), int(ordered[3])))
    tel = np.isin(res["type"], [3, 4])
    got = list(zip(res["i"][tel].tolist(), res["j"][tel].tolist()))
    assert got == sorted(kept) and len(got) > 10
    assert [kept[k] for k in got] == list(zip(res["portal0"].tolist(), res["portal1"].tolist()))
    assert np.any(res["type"] == 4) and np.any(res["type"] == 3)
    assert np.all(np.isnan(res["depth"][tel]))
    assert np.all((res["kick"] != 0.0).any(axis=1) == (res["type"][tel] == 4))


# ---- the product's portal kernels, run on the CPU (tests/portal_kernel_harness.cpp) -----------------------------------
@pytest.fixture(scope="module")
def pk(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("pk") / "libportal_kernels.so")
    src = os.path.join(ROOT, "tests", "portal_kernel_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.pk_set_sort_mode.argtypes = [C.c_int]
    lib.pk_sort.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    lib.pk_set_portals.argtypes = [C.c_uint32] + [C.c_void_p] * 7
    lib.pk_active_set.restype = C.c_int
    lib.pk_active_set.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pk_num_candidates.restype = C.c_uint64
    for f in ("pk_num_boxes", "pk_num_regular", "pk_num_teleported"):
        getattr(lib, f).restype = C.c_uint32
    lib.pk_copy.argtypes = [C.c_void_p] * 14
    lib.pk_enforce.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    return lib


def _pk_run(pk, scene, dx, q0, q1):
    P = scene["portals"]
    a = [np.ascontiguousarray(P[k], dtype=np.float64) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n", "v", "bounds")]
    dx = np.ascontiguousarray(dx, dtype=np.float64)
    pk.pk_set_portals(a[4].shape[0], *[vp(x) for x in a], vp(dx))
    n = scene["r"].shape[0]
    q0, q1, r = [np.ascontiguousarray(x, dtype=np.float64) for x in (q0, q1, scene["r"])]
    if pk.pk_active_set(n, vp(q0), vp(q1), vp(r)) != 0:
        return None
    nc, nb, nr, nt = int(pk.pk_num_candidates()), int(pk.pk_num_boxes()), int(pk.pk_num_regular()), int(pk.pk_num_teleported())
    na = nr + nt
    out = {"candidates": np.zeros((nc, 2), np.uint32), "box_body": np.zeros(nb, np.uint32), "box_portal": np.zeros(nb, np.uint32),
           "type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32), "n": np.zeros((na, 2)), "p": np.zeros((na, 2)),
           "depth": np.zeros(na), "portal0": np.zeros(nt, np.uint32), "portal1": np.zeros(nt, np.uint32), "x0": np.zeros((nt, 2)), "x1": np.zeros((nt, 2)),
           "kick": np.zeros((nt, 2)), "n_regular": nr}
    pk.pk_copy(*[vp(out[k]) for k in ("candidates", "box_body", "box_portal", "type", "i", "j", "n", "p", "depth", "portal0", "portal1", "x0", "x1", "kick")])
    return out


KERNEL_CASES = [dict(n=1, seed=1, axes="xy"), dict(n=2, seed=2, axes="x", side=1.5), dict(n=300, seed=3, axes="xy", side=8.0),
                dict(n=300, seed=4, axes="x", side=8.0, oblique=True), dict(n=1500, seed=5, axes="xy", side=20.0, lees_edwards=0.75, t=3.7),
                dict(n=1500, seed=6, axes="y", side=20.0, lees_edwards=-1.25, t=11.3, oblique=True),
                dict(n=400, seed=11, side=6.0, rmin=0.2, rmax=0.45, axes="xy", static=True)]


@pytest.mark.parametrize("case", KERNEL_CASES, ids=lambda c: "n%d-%s-le%g" % (c["n"], c["axes"], c.get("lees_edwards", 0.0)))
def test_portal_kernels_on_cpu_match_oracle(case, pk, oracle):
    kw = dict(case)
    static = kw.pop("static", False)
    scene = scenes.ball2d_periodic(kw.pop("n"), kw.pop("seed"), **kw)
    o = _oracle(scene)
    dx = o.update_portals(scene["t"])
    q0 = scene["q"]
    q1 = q0.copy() if static else o.flow(0, q0, scene["v"], scene["dt"])[0]
    ref = o.active_set_portals(q0, q1, "allpairs")
    got = _pk_run(pk, scene, dx, q0, q1)
    assert ref is not None and got is not None
    bb = ref["type"] != 2
    for k in ("candidates", "box_body", "box_portal", "portal0", "portal1"):
        assert np.array_equal(got[k], ref[k]), k
    for k in ("type", "i", "j"):
        assert np.array_equal(got[k], ref[k][bb]), k
    for k in ("n", "p", "x0", "x1", "kick"):
        want = ref[k][bb] if k in ("n", "p") else ref[k]
        assert np.array_equal(got[k].view(np.uint64), want.view(np.uint64)), k
    assert np.array_equal(np.isnan(got["depth"]), np.isnan(ref["depth"][bb]))
    ok = ~np.isnan(got["depth"])
    assert np.array_equal(got["depth"][ok], ref["depth"][bb][ok])
    assert got["n_regular"] == ref["n_regular"]


def test_portal_kernels_on_cpu_both_planes_and_enforce(pk, oracle):
    scene = scenes.ball2d_periodic(4, 1, side=1.0, rmin=0.6, rmax=0.7, axes="x")
    assert _pk_run(pk, scene, np.zeros(1), scene["q"], scene["q"]) is None
    scene = scenes.ball2d_periodic(3000, 4, lees_edwards=1.5, t=2.3, oblique=True)
    o = _oracle(scene)
    dx = o.update_portals(scene["t"])
    P = scene["portals"]
    a = [np.ascontiguousarray(P[k], dtype=np.float64) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n", "v", "bounds")]
    pk.pk_set_portals(a[4].shape[0], *[vp(x) for x in a], vp(np.ascontiguousarray(dx)))
    rng = np.random.default_rng(8)
    q = scene["q"] + rng.uniform(-0.45, 0.45, size=scene["q"].shape) * scene["side"]
    rq, rv = o.enforce_portals(q, scene["v"])
    gq, gv = q.copy(), scene["v"].copy()
    pk.pk_enforce(3000, vp(gq), vp(gv))
    assert np.array_equal(gq, rq) and np.array_equal(gv, rv)


@pytest.mark.parametrize("mode,sizes", [(0, (1, 2, 3, 31, 64, 65, 200, 1000, 5000)), (1, (700, 2048, 5000))], ids=["tile64", "tile2048"])
def test_tile_sort_kernel_on_cpu(pk, mode, sizes):
    """k_b2p_bitonic_tile (+ k_b2p_bitonic for the strides that cross tiles), launch sequence as in the driver: one tile,
    several tiles, lists shorter than a tile; mode 1 uses the library's own tile size and thread count."""
    pk.pk_set_sort_mode(mode)
    rng = np.random.default_rng(5)
    try:
        for nraw in sizes:
            b0 = rng.integers(0, max(2, nraw // 3), size=nraw).astype(np.uint64)
            b1 = b0 + rng.integers(1, 4, size=nraw).astype(np.uint64)
            keys = (b0 << np.uint64(32)) | b1
            m = 1
            while m < nraw:
                m <<= 1
            k = np.full(m, np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
            i = np.full(m, 0xFFFFFFFF, dtype=np.uint32)
            k[:nraw] = keys
            i[:nraw] = np.arange(nraw, dtype=np.uint32)
            pk.pk_sort(m, vp(k), vp(i))
            order = np.lexsort((np.arange(nraw), keys))
            assert np.array_equal(k[:nraw], keys[order]) and np.array_equal(i[:nraw], order.astype(np.uint32)), nraw
            assert np.all(k[nraw:] == np.uint64(0xFFFFFFFFFFFFFFFF))
    finally:
        pk.pk_set_sort_mode(0)


# ---- rigidbody2d portals ------------------------------------------------------------------------------------------------------
def _rb2d_oracle(scene):
    o = ob.RB2DOracle(scene)
    o.set_portals(scene["portals"])
    return o


def _bind_rb2d(pk):
    if getattr(pk, "_rb2d_bound", False):
        return
    pk.pk_rb2d_set_portals.argtypes = [C.c_uint32] + [C.c_void_p] * 7
    pk.pk_rb2d_active_set.restype = C.c_uint
    pk.pk_rb2d_active_set.argtypes = [C.c_uint32] + [C.c_void_p] * 4
    for f in ("pk_rb2d_num_candidates", "pk_rb2d_num_regular_pairs"):
        getattr(pk, f).restype = C.c_uint64
    for f in ("pk_rb2d_num_boxes", "pk_rb2d_num_teleported"):
        getattr(pk, f).restype = C.c_uint32
    pk.pk_rb2d_copy.argtypes = [C.c_void_p] * 17
    pk.pk_rb2d_enforce.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    pk.pk_rb2d_probe.restype = C.c_uint32
    pk.pk_rb2d_probe.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    pk._rb2d_bound = True


def _pk_rb2d_set(pk, scene, dx):
    _bind_rb2d(pk)
    P = scene["portals"]
    a = [np.ascontiguousarray(P[k], dtype=np.float64) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n", "v", "bounds")]
    pk.pk_rb2d_set_portals(a[4].shape[0], *[vp(x) for x in a], vp(np.ascontiguousarray(dx, dtype=np.float64)))


def _rb2d_device_arrays(scene):
    """btype / bparam exactly as sg_rb2d_set_bodies lays them out."""
    gi = scene["geo_of_body"]
    t = scene["geo_type"][gi].astype(np.uint32)
    btype = (t | (scene["fixed"].astype(np.uint32) << np.uint32(31))).astype(np.uint32)
    bparam = np.zeros((gi.shape[0], 2))
    circ = t == 0
    bparam[circ, 0] = scene["geo_r"][gi][circ]
    bparam[~circ] = scene["geo_half"][gi][~circ]
    return np.ascontiguousarray(btype), np.ascontiguousarray(bparam)


@pytest.mark.parametrize("case", [dict(axes="xy"), dict(axes="xy", lees_edwards=0.75, t=3.7), dict(axes="x", lees_edwards=-1.25, t=11.3, oblique=True)],
                         ids=["plain", "lees-edwards", "oblique"])
def test_rb2d_portal_primitives_oracle_reference_product(case, pk, oracle):
    scene = scenes.rb2d_periodic(32, 5, side=12.0, **case)
    o = _rb2d_oracle(scene)
    dx = o.update_portals(scene["t"])
    _pk_rb2d_set(pk, scene, dx)
    path = os.path.join(REFDIR, "libref_rb2d.so")
    ref, handles = None, []
    if os.path.exists(path) and hasattr(C.CDLL(path), "ref_rb2d_portal_probe"):
        ref = C.CDLL(path)
        ref.ref_rb2d_portal_create.restype = C.c_void_p
        ref.ref_rb2d_portal_create.argtypes = [C.c_void_p] * 4 + [C.c_double, C.c_double]
        ref.ref_rb2d_portal_update.argtypes = [C.c_void_p, C.c_double]
        ref.ref_rb2d_portal_probe.restype = C.c_uint32
        ref.ref_rb2d_portal_probe.argtypes = [C.c_void_p] * 4
        ref.ref_rb2d_portal_destroy.argtypes = [C.c_void_p]
        P = scene["portals"]
        for p in range(P["v"].shape[0]):
            h = ref.ref_rb2d_portal_create(vp(P["plane_a_x"][p]), vp(P["plane_a_n"][p]), vp(P["plane_b_x"][p]), vp(P["plane_b_n"][p]), float(P["v"][p]), float(P["bounds"][p]))
            ref.ref_rb2d_portal_update(h, scene["t"])
            handles.append(h)
    elif os.path.isdir("/root/reference"):
        pytest.fail("oracle/_ref/libref_rb2d.so lacks the portal shim: run make -C oracle -f Makefile.ref")
    rng = np.random.default_rng(23)
    ctr = _probe_points(scene, rng, 3000)
    ext = rng.uniform(0.05, 0.6, size=(ctr.shape[0], 2))
    seen = set()
    for p in range(scene["portals"]["v"].shape[0]):
        for c, e in zip(ctr, ext):
            box = np.ascontiguousarray(np.concatenate([c - e, c + e]))
            c = np.ascontiguousarray(c)
            to, oo = o.portal_probe(p, box, c)
            op = np.zeros(4)
            tp = pk.pk_rb2d_probe(p, vp(box), vp(c), vp(op))
            assert to == tp and np.array_equal(oo.view(np.uint64), op.view(np.uint64)), (p, box)
            if ref is not None:
                orf = np.zeros(4)
                tr = ref.ref_rb2d_portal_probe(handles[p], vp(box), vp(c), vp(orf))
                assert tr == to and np.array_equal(orf.view(np.uint64), oo.view(np.uint64)), (p, box)
            seen.add(to)
    assert seen == {0, 1, 2}
    for h in handles:
        ref.ref_rb2d_portal_destroy(h)


RB2D_KERNEL_CASES = [dict(n=1, seed=1), dict(n=2, seed=2, side=3.0, axes="x"), dict(n=400, seed=1, side=10.0), dict(n=400, seed=2, side=10.0, lees_edwards=0.7, t=1.3, oblique=True),
                     dict(n=600, seed=3, side=16.0, boxes=True, axes="x"), dict(n=400, seed=4, side=10.0, nfixed_frac=0.3), dict(n=500, seed=6, side=7.0, axes="y", lees_edwards=-0.9, t=4.0)]


@pytest.mark.parametrize("case", RB2D_KERNEL_CASES, ids=lambda c: "n%d-s%d" % (c["n"], c["seed"]))
def test_rb2d_portal_kernels_on_cpu_match_oracle(case, pk, oracle):
    scene = scenes.rb2d_periodic(**case)
    o = _rb2d_oracle(scene)
    dx = o.update_portals(scene["t"])
    q0 = scene["q"]
    q1, _ = o.flow(0, q0, scene["v"], scene["dt"])
    ref = o.active_set_portals(q0, q1, "allpairs")
    assert ref["supported"]
    _pk_rb2d_set(pk, scene, dx)
    btype, bparam = _rb2d_device_arrays(scene)
    q0c, q1c = np.ascontiguousarray(q0), np.ascontiguousarray(q1)
    bad = pk.pk_rb2d_active_set(scene["geo_of_body"].shape[0], vp(btype), vp(bparam), vp(q0c), vp(q1c))
    assert bad == 0
    nc, nrp, nb, nt = int(pk.pk_rb2d_num_candidates()), int(pk.pk_rb2d_num_regular_pairs()), int(pk.pk_rb2d_num_boxes()), int(pk.pk_rb2d_num_teleported())
    got = {"candidates": np.zeros((nc, 2), np.uint32), "reg_pairs": np.zeros((nrp, 2), np.uint32), "box_body": np.zeros(nb, np.uint32), "box_portal": np.zeros(nb, np.uint32),
           "type": np.zeros(nt, np.uint32), "i": np.zeros(nt, np.uint32), "j": np.zeros(nt, np.uint32), "n": np.zeros((nt, 2)), "p": np.zeros((nt, 2)), "depth": np.zeros(nt),
           "portal0": np.zeros(nt, np.uint32), "portal1": np.zeros(nt, np.uint32), "x0": np.zeros((nt, 2)), "x1": np.zeros((nt, 2)), "delta0": np.zeros((nt, 2)), "delta1": np.zeros((nt, 2)),
           "kick": np.zeros((nt, 2))}
    pk.pk_rb2d_copy(*[vp(got[k]) for k in ("candidates", "reg_pairs", "box_body", "box_portal", "type", "i", "j", "n", "p", "depth", "portal0", "portal1", "x0", "x1", "delta0", "delta1", "kick")])
    n = scene["geo_of_body"].shape[0]
    for k in ("candidates", "box_body", "box_portal", "portal0", "portal1"):
        assert np.array_equal(got[k], ref[k]), k
    real = (ref["candidates"][:, 0] < n) & (ref["candidates"][:, 1] < n)
    assert np.array_equal(got["reg_pairs"], ref["candidates"][real])
    nr = ref["n_regular"]
    tel = slice(nr, nr + nt)
    assert np.all(np.isin(ref["type"][tel], [25, 26])) and not np.any(np.isin(ref["type"][nr + nt:], [25, 26]))
    for k in ("type", "i", "j"):
        assert np.array_equal(got[k], ref[k][tel]), k
    eq = lambda a, b: np.array_equal(a.view(np.uint64), b.view(np.uint64)) or (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]))
    for k in ("n", "p", "depth"):
        assert eq(got[k], ref[k][tel]), k
    for k in ("x0", "x1", "delta0", "delta1", "kick"):
        assert eq(got[k], ref[k]), k
    if case["n"] >= 400:
        assert nt > 5
    if case.get("lees_edwards", 0.0) != 0.0:
        assert np.any(got["type"] == 26) and np.all(np.isnan(got["delta0"][got["type"] == 26])) and np.all(got["depth"][got["type"] == 26] == 0.0)


def test_rb2d_portal_kernels_on_cpu_unsupported_and_enforce(pk, oracle):
    # boxes at the portals: the reference exits, the classify kernel raises bit 0
    scene = scenes.rb2d_periodic(300, 8, side=9.0, boxes=True)
    rng = np.random.default_rng(2)
    q = scene["q"].reshape(-1, 3)
    q[:, :2] = rng.uniform(0.0, scene["side"], size=q[:, :2].shape)
    scene["q"] = q.ravel().copy()
    o = _rb2d_oracle(scene)
    dx = o.update_portals(0.0)
    assert not o.active_set_portals(scene["q"], scene["q"], "allpairs")["supported"]
    _pk_rb2d_set(pk, scene, dx)
    btype, bparam = _rb2d_device_arrays(scene)
    qc = np.ascontiguousarray(scene["q"])
    assert pk.pk_rb2d_active_set(300, vp(btype), vp(bparam), vp(qc), vp(qc)) & 1
    # kinematic circles at the portals: bit 1
    scene = scenes.rb2d_periodic(300, 9, side=9.0)
    scene["fixed"][::3] = 1
    o = _rb2d_oracle(scene)
    dx = o.update_portals(0.0)
    assert not o.active_set_portals(scene["q"], scene["q"], "allpairs")["supported"]
    _pk_rb2d_set(pk, scene, dx)
    btype, bparam = _rb2d_device_arrays(scene)
    qc = np.ascontiguousarray(scene["q"])
    assert pk.pk_rb2d_active_set(300, vp(btype), vp(bparam), vp(qc), vp(qc)) & 2
    # enforce
    scene = scenes.rb2d_periodic(2000, 4, lees_edwards=1.5, t=2.3, oblique=True)
    o = _rb2d_oracle(scene)
    dx = o.update_portals(scene["t"])
    _pk_rb2d_set(pk, scene, dx)
    q = scene["q"].copy()
    q.reshape(-1, 3)[:, :2] += np.random.default_rng(8).uniform(-0.45, 0.45, size=(2000, 2)) * scene["side"]
    rq, rv = o.enforce_portals(q, scene["v"])
    gq, gv = q.copy(), scene["v"].copy()
    pk.pk_rb2d_enforce(2000, vp(gq), vp(gv))
    assert np.array_equal(gq, rq) and np.array_equal(gv, rv) and np.any(gq != q) and np.any(gv != scene["v"])
