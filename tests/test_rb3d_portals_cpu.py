"""CPU: rigidbody3d portals, oracle first (oracle/rb3d_portals.h; the GPU side of this row is not built yet).

  * StaticPlane frames ( n, t0, t1 ) and the PlanarPortal primitives ( box touch test, teleports through A / B, point inside,
    teleportPointInsidePortal, integer portal multipliers ): restatement == the reference's own rigidbody3d/StaticGeometry/
    StaticPlane.cpp + rigidbody3d/Portals/PlanarPortal.cpp compiled unchanged (oracle/_ref), bit for bit
  * with no portals the portal path returns exactly the regular active set (it IS the body-body path in the reference)
  * periodic box: the oracle's sphere contacts == brute-force minimum-image contacts (exact arithmetic on a dyadic grid)
  * order, kinematic variants, enforcePeriodicBoundaryConditions
"""
import ctypes as C
import os

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def _ref():
    path = os.path.join(REFDIR, "libref_rb3d.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_rb3d.so not built (no /root/reference in this container)")
    lib = C.CDLL(path)
    if not hasattr(lib, "ref_rb3d_portal_probe"):
        if os.path.isdir("/root/reference"):
            pytest.fail("oracle/_ref/libref_rb3d.so lacks the portal shim: run make -C oracle -f Makefile.ref")
        pytest.skip("oracle/_ref predates the portal shim")
    lib.ref_rb3d_plane_frame.argtypes = [C.c_void_p] * 3
    lib.ref_rb3d_portal_create.restype = C.c_void_p
    lib.ref_rb3d_portal_create.argtypes = [C.c_void_p] * 5
    lib.ref_rb3d_portal_probe.restype = C.c_uint32
    lib.ref_rb3d_portal_probe.argtypes = [C.c_void_p] * 4
    lib.ref_rb3d_portal_destroy.argtypes = [C.c_void_p]
    return lib


def _oracle(scene):
    o = ob.RB3DOracle(scene)
    o.set_portals(scene["portals"])
    return o


def test_plane_frames_match_the_reference(oracle):
    ref = _ref()
    o = ob.RB3DOracle(scenes.rb3d_random_spheres(2, 1))
    rng = np.random.default_rng(3)
    normals = rng.normal(size=(3000, 3)) * rng.uniform(0.2, 4.0, size=(3000, 1))
    normals[:6] = [[0, 1, 0], [1, 0, 0], [0, 0, 1], [-1, 0, 0], [0, 0, -1], [0, 2.5, 0]]
    normals = normals[normals[:, 1] / np.linalg.norm(normals, axis=1) > -0.999]  # Eigen's SVD branch is not restated
    for n in normals:
        x = np.ascontiguousarray(rng.uniform(-5, 5, size=3))
        n = np.ascontiguousarray(n)
        a = o.plane_frame(x, n)
        b = np.zeros(9)
        ref.ref_rb3d_plane_frame(vp(x), vp(n), vp(b))
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), n
        f = a.reshape(3, 3)
        assert np.allclose(f @ f.T, np.eye(3), atol=1e-12)


@pytest.mark.parametrize("case", [dict(axes="xz"), dict(axes="xyz", tilt=True), dict(axes="z", mult=(1, -1, 1), tilt=True)], ids=["xz", "xyz-tilt", "mult"])
def test_portal_primitives_match_the_reference(case, oracle):
    ref = _ref()
    scene = scenes.rb3d_periodic_spheres(8, 2, side=9.0, **case)
    o = _oracle(scene)
    P = scene["portals"]
    handles = [ref.ref_rb3d_portal_create(vp(P["plane_a_x"][p]), vp(P["plane_a_n"][p]), vp(P["plane_b_x"][p]), vp(P["plane_b_n"][p]), vp(P["mult"][p])) for p in range(P["mult"].shape[0])]
    rng = np.random.default_rng(11)
    side = scene["side"]
    seen = set()
    for p, h in enumerate(handles):
        for k in range(2500):
            c = rng.uniform(-0.4 * side, 1.4 * side, size=3)
            if k % 3 == 0:
                c[rng.integers(0, 3)] = [0.0, side][int(rng.integers(0, 2))] + rng.uniform(-0.5, 0.5)
            e = rng.uniform(0.05, 0.7, size=3)
            box = np.ascontiguousarray(np.concatenate([c - e, c + e]))
            c = np.ascontiguousarray(c)
            co, oo = o.portal_probe(p, box, c)
            orf = np.zeros(9)
            cr = ref.ref_rb3d_portal_probe(h, vp(box), vp(c), vp(orf))
            assert co == cr and np.array_equal(oo.view(np.uint64), orf.view(np.uint64)), (p, box)
            seen.add(co & 3)
    assert seen == {0, 1, 2}
    for h in handles:
        ref.ref_rb3d_portal_destroy(h)


def test_no_portals_is_the_regular_path(oracle):
    for s in (scenes.rb3d_random_spheres(500, 2, nfixed_frac=0.3), scenes.rb3d_random_boxes(150, 3), scenes.rb3d_mixed_segregated(40, 5)):
        o = ob.RB3DOracle(s)
        o.set_portals({"plane_a_x": np.zeros((0, 3)), "plane_a_n": np.zeros((0, 3)), "plane_b_x": np.zeros((0, 3)), "plane_b_n": np.zeros((0, 3)), "mult": np.zeros((0, 3), np.int32)})
        q1, _ = o.flow(2, s["q"], s["v"], s["dt"])
        a = o.active_set(s["q"], q1, "allpairs")
        b = o.active_set_portals(s["q"], q1, "allpairs")
        assert a["supported"] and b["supported"]
        for k in ("candidates", "type", "i", "j", "aux"):
            assert np.array_equal(a[k], b[k]), k
        for k in ("n", "p", "depth"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k
        assert b["n_regular"] == int(np.isin(b["type"], [10, 11, 12, 13]).sum())


def _min_image_pairs(x, r, side, axes):
    d = x[:, None, :] - x[None, :, :]
    for k, ax in enumerate("xyz"):
        if ax in axes:
            d[..., k] -= side * np.round(d[..., k] / side)
    hit = (d ** 2).sum(-1) <= (r[:, None] + r[None, :]) ** 2
    i, j = np.nonzero(np.triu(hit, 1))
    return set(zip(i.tolist(), j.tolist()))


@pytest.mark.parametrize("axes", ["x", "z"])
def test_oracle_active_set_equals_minimum_image(axes, oracle):
    n, side = 600, 8.0
    scene = scenes.rb3d_periodic_spheres(n, 4, side=side, axes=axes)
    rng = np.random.default_rng(9)
    x = np.round(rng.uniform(0.0, side, size=(n, 3)) * 1024.0) / 1024.0
    x = np.minimum(x, side - 1.0 / 1024.0)
    scene["q"][: 3 * n] = x.ravel()
    scene["geo_r"] = np.array([0.25, 0.375, 0.5])
    P = scene["portals"]
    P["plane_a_n"], P["plane_b_n"] = np.sign(P["plane_a_n"]), np.sign(P["plane_b_n"])
    o = _oracle(scene)
    res = o.active_set_portals(scene["q"], scene["q"], "allpairs")
    assert res["supported"]
    bb = np.isin(res["type"], [10, 19])
    got = set(zip(np.minimum(res["i"][bb], res["j"][bb]).tolist(), np.maximum(res["i"][bb], res["j"][bb]).tolist()))
    assert len(got) == int(bb.sum())
    r = scene["geo_r"][scene["geo_of_body"]]
    assert got == _min_image_pairs(x, r, side, axes)
    nr, nt = res["n_regular"], res["portal0"].shape[0]
    assert nt > 5 and np.all(res["type"][:nr] == 10) and np.all(res["type"][nr:nr + nt] == 19) and not np.any(np.isin(res["type"][nr + nt:], [10, 19]))
    for blk in (slice(0, nr), slice(nr, nr + nt)):
        key = res["i"][blk].astype(np.int64) * n + res["j"][blk]
        assert np.all(np.diff(key) > 0)
    assert np.all(np.isnan(res["depth"][nr:nr + nt]))
    res2 = o.active_set_portals(scene["q"], scene["q"], "grid")
    assert np.array_equal(res["candidates"], res2["candidates"]) and np.array_equal(res["i"], res2["i"])


def test_kinematic_variants_and_enforce(oracle):
    scene = scenes.rb3d_periodic_spheres(700, 6, side=7.0, axes="xyz", nfixed_frac=0.35, tilt=True)
    o = _oracle(scene)
    q1, _ = o.flow(2, scene["q"], scene["v"], scene["dt"])
    res = o.active_set_portals(scene["q"], q1, "allpairs")
    assert res["supported"]
    nr, nt = res["n_regular"], res["portal0"].shape[0]
    tel = res["type"][nr:nr + nt]
    assert (tel == 19).sum() > 3 and (tel == 30).sum() > 3
    fixed = scene["fixed"].astype(bool)
    kin = tel == 30
    i, j = res["i"][nr:nr + nt], res["j"][nr:nr + nt]
    assert np.all(~fixed[i[kin]]) and np.all(fixed[j[kin]])          # free sphere first, kinematic object second
    assert np.all(~fixed[i[~kin]]) and np.all(~fixed[j[~kin]]) and np.all(i[~kin] < j[~kin])
    # the kinematic-object contact carries the kinematic body's teleported centre
    x0, x1 = res["x0"], res["x1"]
    swapped = kin & (i > j)
    assert np.array_equal(res["p"][nr:nr + nt][kin & ~swapped], x1[kin & ~swapped]) and np.array_equal(res["p"][nr:nr + nt][swapped], x0[swapped])
    assert np.allclose(np.linalg.norm(res["n"][nr:nr + nt], axis=1), 1.0)
    # enforce: afterwards nobody is inside a portal; rotations untouched
    n = scene["geo_of_body"].shape[0]
    q = scene["q"].copy()
    q[: 3 * n] += np.random.default_rng(5).uniform(-0.4, 0.4, size=3 * n) * scene["side"]
    q2 = o.enforce_portals(q)
    assert np.array_equal(q2[3 * n:], q[3 * n:]) and np.any(q2[: 3 * n] != q[: 3 * n])
    for p in range(scene["portals"]["mult"].shape[0]):
        for xb in q2[: 3 * n].reshape(-1, 3)[::5]:
            code, _ = o.portal_probe(p, np.zeros(6), np.ascontiguousarray(xb))
            assert not (code & 4)


# ---- the product's rigidbody3d portal kernels, run on the CPU (tests/portal_kernel_harness.cpp) ------------------------------
@pytest.fixture(scope="module")
def pk(tmp_path_factory):
    import subprocess
    out = str(tmp_path_factory.mktemp("pk3") / "libportal_kernels.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-o", out, os.path.join(ROOT, "tests", "portal_kernel_harness.cpp")], check=True)
    lib = C.CDLL(out)
    lib.pk_rb3d_set_portals.restype = C.c_int
    lib.pk_rb3d_set_portals.argtypes = [C.c_uint32] + [C.c_void_p] * 5
    lib.pk_rb3d_active_set.argtypes = [C.c_uint32] + [C.c_void_p] * 4
    for f in ("pk_rb3d_num_candidates", "pk_rb3d_num_regular_pairs"):
        getattr(lib, f).restype = C.c_uint64
    for f in ("pk_rb3d_num_boxes", "pk_rb3d_num_teleported"):
        getattr(lib, f).restype = C.c_uint32
    lib.pk_rb3d_copy.argtypes = [C.c_void_p] * 14
    lib.pk_rb3d_enforce.argtypes = [C.c_uint32, C.c_void_p]
    return lib


def _pk3_set(pk, scene):
    P = scene["portals"]
    a = [np.ascontiguousarray(P[k], dtype=np.float64) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n")]
    m = np.ascontiguousarray(P["mult"], dtype=np.int32)
    assert pk.pk_rb3d_set_portals(m.shape[0], *[vp(x) for x in a], vp(m)) == 1


K3_CASES = [dict(n=1, seed=1), dict(n=2, seed=2, side=3.0, axes="x"), dict(n=500, seed=3, side=7.0, axes="xz"), dict(n=600, seed=4, side=7.0, axes="xyz", tilt=True),
            dict(n=700, seed=6, side=7.0, axes="xyz", nfixed_frac=0.35, tilt=True), dict(n=400, seed=7, side=5.0, axes="z", mult=(1, 1, 1))]


@pytest.mark.parametrize("case", K3_CASES, ids=lambda c: "n%d-s%d" % (c["n"], c["seed"]))
def test_rb3d_portal_kernels_on_cpu_match_oracle(case, pk, oracle):
    scene = scenes.rb3d_periodic_spheres(**case)
    o = _oracle(scene)
    q0 = scene["q"]
    q1, _ = o.flow(2, q0, scene["v"], scene["dt"])
    ref = o.active_set_portals(q0, q1, "allpairs")
    assert ref["supported"]
    _pk3_set(pk, scene)
    n = scene["geo_of_body"].shape[0]
    btype = np.ascontiguousarray(np.uint32(1) | (scene["fixed"].astype(np.uint32) << np.uint32(31)))
    bparam = np.zeros((n, 4))
    bparam[:, 0] = scene["geo_r"][scene["geo_of_body"]]
    q0c, q1c = np.ascontiguousarray(q0), np.ascontiguousarray(q1)
    pk.pk_rb3d_active_set(n, vp(btype), vp(bparam), vp(q0c), vp(q1c))
    nc, nrp, nb, nt = int(pk.pk_rb3d_num_candidates()), int(pk.pk_rb3d_num_regular_pairs()), int(pk.pk_rb3d_num_boxes()), int(pk.pk_rb3d_num_teleported())
    got = {"candidates": np.zeros((nc, 2), np.uint32), "reg_pairs": np.zeros((nrp, 2), np.uint32), "box_body": np.zeros(nb, np.uint32), "box_portal": np.zeros(nb, np.uint32),
           "type": np.zeros(nt, np.uint32), "i": np.zeros(nt, np.uint32), "j": np.zeros(nt, np.uint32), "n": np.zeros((nt, 3)), "p": np.zeros((nt, 3)), "depth": np.zeros(nt),
           "portal0": np.zeros(nt, np.uint32), "portal1": np.zeros(nt, np.uint32), "x0": np.zeros((nt, 3)), "x1": np.zeros((nt, 3))}
    pk.pk_rb3d_copy(*[vp(got[k]) for k in ("candidates", "reg_pairs", "box_body", "box_portal", "type", "i", "j", "n", "p", "depth", "portal0", "portal1", "x0", "x1")])
    for k in ("candidates", "box_body", "box_portal", "portal0", "portal1"):
        assert np.array_equal(got[k], ref[k]), k
    real = (ref["candidates"][:, 0] < n) & (ref["candidates"][:, 1] < n)
    assert np.array_equal(got["reg_pairs"], ref["candidates"][real])
    nr = ref["n_regular"]
    tel = slice(nr, nr + nt)
    assert nt == ref["portal0"].shape[0]
    for k in ("type", "i", "j"):
        assert np.array_equal(got[k], ref[k][tel]), k
    for k in ("n", "p"):
        assert np.array_equal(got[k].view(np.uint64), ref[k][tel].view(np.uint64)), k
    assert np.all(np.isnan(got["depth"]))
    for k in ("x0", "x1"):
        assert np.array_equal(got[k].view(np.uint64), ref[k].view(np.uint64)), k
    if case["n"] >= 400:
        assert nt > 3
    if case.get("nfixed_frac", 0.0) > 0.0:
        assert np.any(got["type"] == 30) and np.any(got["type"] == 19)


def test_rb3d_portal_enforce_kernel_on_cpu(pk, oracle):
    scene = scenes.rb3d_periodic_spheres(3000, 8, axes="xyz", tilt=True)
    o = _oracle(scene)
    _pk3_set(pk, scene)
    n = 3000
    q = scene["q"].copy()
    q[: 3 * n] += np.random.default_rng(5).uniform(-0.4, 0.4, size=3 * n) * scene["side"]
    ref = o.enforce_portals(q)
    got = q.copy()
    pk.pk_rb3d_enforce(n, vp(got))
    assert np.array_equal(got, ref) and np.any(got != q)
