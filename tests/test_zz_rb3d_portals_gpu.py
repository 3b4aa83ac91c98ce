"""GPU parity (pytest -m gpu): rigidbody3d planar portals (all-sphere scenes) through the C ABI vs the CPU oracle
(oracle/rb3d_portals.h; its plane frames and portal primitives are checked against the reference's compiled StaticPlane.cpp /
PlanarPortal.cpp, and the kernels themselves run on the CPU in tests/test_rb3d_portals_cpu.py).

Like tests/test_zz_rb2d_portals_gpu.py this was written after round 1's GPU budget was spent: kernels verified in emulation, the
host driver (the rigidbody2d portal driver with k_rb3d_pairs spliced in) first executes when this file runs; it sorts last.

Bar: extended candidate list, teleported-box table, active set in the reference's order (contacts of un-teleported pairs |
teleported | planes) and the teleported centres bit-identical.
"""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu


def make_sim(s, ctx):
    import scisim_b200 as sb
    st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"],
                             planar_portals=sb.PlanarPortal3D.from_arrays(s["portals"]))
    return sb.RigidBody3DSim(st, ctx=ctx)


def make_oracle(s):
    from tests import oracle_binding as ob
    o = ob.RB3DOracle(s)
    o.set_portals(s["portals"])
    return o


def assert_equal(gpu, tele, ref):
    assert ref["supported"]
    assert gpu.n_candidates == ref["candidates"].shape[0] and np.array_equal(gpu.candidates, ref["candidates"])
    assert np.array_equal(tele.box_body, ref["box_body"]) and np.array_equal(tele.box_portal, ref["box_portal"])
    assert gpu.n_active == ref["type"].shape[0]
    for k in ("type", "i", "j", "aux"):
        assert np.array_equal(getattr(gpu, k), ref[k]), k
    for k in ("n", "p", "depth"):
        assert np.array_equal(getattr(gpu, k), ref[k], equal_nan=True), k
    assert tele.n_regular == ref["n_regular"] and tele.n_teleported == ref["portal0"].shape[0]
    assert gpu.n_body_body == tele.n_regular + tele.n_teleported
    assert np.array_equal(tele.portal0, ref["portal0"]) and np.array_equal(tele.portal1, ref["portal1"])
    assert np.array_equal(tele.x0, ref["x0"]) and np.array_equal(tele.x1, ref["x1"])
    assert tele.kick is None and tele.delta0 is None


CASES = [dict(n=1, seed=1), dict(n=2, seed=2, side=3.0, axes="x"), dict(n=500, seed=3, side=7.0, axes="xz"), dict(n=600, seed=4, side=7.0, axes="xyz", tilt=True),
         dict(n=700, seed=6, side=7.0, axes="xyz", nfixed_frac=0.35, tilt=True), dict(n=400, seed=7, side=5.0, axes="z", mult=(1, 1, 1)), dict(n=30000, seed=9, axes="xz")]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d-s%d" % (c["n"], c["seed"]))
def test_rb3d_portal_active_set_matches_oracle(gpu_ctx, oracle, case):
    s = scenes.rb3d_periodic_spheres(**case)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    q1, _ = o.flow(2, s["q"], s["v"], s["dt"])
    ref = o.active_set_portals(s["q"], q1, "grid" if case["n"] > 3000 else "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    tele = sim.teleported()
    assert_equal(got, tele, ref)
    if case["n"] >= 400:
        assert tele.n_teleported > 3
    if case.get("nfixed_frac", 0.0) > 0.0:
        assert np.any(got.type == 30) and np.any(got.type == 19)


def test_rb3d_portal_resident_step_and_enforce(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb3d_periodic_spheres(4000, 13, axes="xyz", tilt=True)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    sim.upload(s["q"], s["v"])
    nc, na = sim.step(sb.DMVMap(), s["dt"])
    q1, v1, got = sim.fetch()
    rq1, rv1 = o.flow(3, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    ref = o.active_set_portals(s["q"], rq1)
    assert nc == ref["candidates"].shape[0] and na == ref["type"].shape[0]
    assert_equal(got, sim.teleported(), ref)
    n = 4000
    q = s["q"].copy()
    q[: 3 * n] += np.random.default_rng(5).uniform(-0.4, 0.4, size=3 * n) * s["side"]
    rq = o.enforce_portals(q)
    gq = sim.enforcePeriodicBoundaryConditions(q)
    assert np.array_equal(gq, rq) and np.any(gq != q)


def test_rb3d_portal_limits(gpu_ctx, oracle):
    import scisim_b200 as sb
    # boxes in a scene with portals: refused (the reference's teleported collisions are sphere-only)
    s = scenes.rb3d_random_boxes(50, 3, nplanes=0)
    p = scenes.rb3d_periodic_spheres(4, 1, side=40.0)["portals"]
    st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"],
                             planar_portals=sb.PlanarPortal3D.from_arrays(p))
    sim = sb.RigidBody3DSim(st, ctx=gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"])
    # a plane facing -y: Eigen's SVD branch, refused when the portals are set
    bad = {k: v.copy() for k, v in p.items()}
    bad["plane_a_n"][0] = [0.0, -1.0, 0.0]
    st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], planar_portals=sb.PlanarPortal3D.from_arrays(bad))
    with pytest.raises(sb.SciSimB200Error):
        sb.RigidBody3DSim(st, ctx=gpu_ctx)


def test_rb3d_portals_cleared_restores_the_fused_sphere_path(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    sp = scenes.rb3d_periodic_spheres(300, 3, side=6.0)
    make_sim(sp, gpu_ctx).computeActiveSet(sp["q"], sp["q"])
    s = scenes.rb3d_random_spheres(2000, 4)
    st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"])
    sim = sb.RigidBody3DSim(st, ctx=gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(2, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    assert np.array_equal(got.candidates, ref["candidates"]) and np.array_equal(got.type, ref["type"])
    assert np.array_equal(got.i, ref["i"]) and np.array_equal(got.j, ref["j"]) and np.array_equal(got.n, ref["n"])
    with pytest.raises(sb.SciSimB200Error):
        sim.teleported()


def test_rb3d_portal_trajectory(gpu_ctx, oracle):
    """RigidBody3DSim::flow's portal bookkeeping over 10 steps without contact response (RigidBody3DSim.cpp:513-550)."""
    import scisim_b200 as sb
    s = scenes.rb3d_periodic_spheres(2000, 21, side=14.0, axes="xz")
    s["v"][: 3 * 2000] *= 8.0
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    rq, rv = q.copy(), v.copy()
    crossed = 0
    for it in range(1, 11):
        q1, v1 = sb.SplitHamMap().flow(q, v, sim, it, s["dt"])
        rq1, rv1 = o.flow(2, rq, rv, s["dt"])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
        ref = o.active_set_portals(rq, rq1)
        got = sim.computeActiveSet(q, q1, resident=True)
        assert_equal(got, sim.teleported(), ref)
        q = sim.enforcePeriodicBoundaryConditions(q1)
        rq = o.enforce_portals(rq1)
        v, rv = v1, rv1
        assert np.array_equal(q, rq)
        crossed += int(np.any(q[: 6000].reshape(-1, 3) != q1[: 6000].reshape(-1, 3), axis=1).sum())
    assert crossed > 30
