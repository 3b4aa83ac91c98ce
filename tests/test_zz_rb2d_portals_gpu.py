"""GPU parity (pytest -m gpu): rigidbody2d with planar / Lees-Edwards portals through the C ABI vs the CPU oracle
(oracle/rb2d_portals.h; its portal primitives are checked against the reference's rigidbody2d/PlanarPortal.cpp and the
kernels themselves run on the CPU in tests/test_portals_cpu.py).

Written after round 1's GPU budget was spent: the kernels are verified in emulation, the host driver's launch sequence
(modelled on the ball2d portal driver, which passed on the B200) first executes when this file runs -- it sorts last so that
a problem here cannot hide the rest of the suite.

Bar: extended candidate list, teleported-box table, the active set in the reference's order (contacts of un-teleported pairs |
teleported | planes) and the constructor arguments of the teleported contacts bit-identical; contacts that involve rotated
boxes within 1e-12 (sincos), as in tests/test_rb2d_gpu.py.
"""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
REL = 1.0e-12


def make_sim(scene, ctx):
    import scisim_b200 as sb
    pp = sb.PlanarPortal.from_arrays(scene["portals"])
    st = sb.RigidBody2DState(scene["geo_type"], scene["geo_r"], scene["geo_half"], scene["geo_of_body"], scene["fixed"], scene["M"], scene["g"], scene["plane_x"], scene["plane_n"],
                             planar_portals=pp)
    return sb.RigidBody2DSim(st, ctx=ctx)


def make_oracle(scene):
    from tests import oracle_binding as ob
    o = ob.RB2DOracle(scene)
    o.set_portals(scene["portals"])
    return o


def close(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.all(np.abs(a[~np.isnan(a)] - b[~np.isnan(b)]) <= REL * np.maximum(1.0, np.abs(b[~np.isnan(b)])))


def assert_equal(gpu, tele, ref, exact):
    assert ref["supported"]
    assert gpu.n_candidates == ref["candidates"].shape[0] and np.array_equal(gpu.candidates, ref["candidates"])
    assert np.array_equal(tele.box_body, ref["box_body"]) and np.array_equal(tele.box_portal, ref["box_portal"])
    assert gpu.n_active == ref["type"].shape[0]
    for k in ("type", "i", "j", "aux"):
        assert np.array_equal(getattr(gpu, k), ref[k]), k
    for k in ("n", "p", "depth"):
        assert close(getattr(gpu, k), ref[k]), k
        if exact:
            g, r = getattr(gpu, k), ref[k]
            assert np.array_equal(g[~np.isnan(g)], r[~np.isnan(r)]), k + " within tolerance but not bit-identical"
    assert tele.n_regular == ref["n_regular"] and tele.n_teleported == ref["portal0"].shape[0]
    assert gpu.n_body_body == tele.n_regular + tele.n_teleported
    assert np.array_equal(tele.portal0, ref["portal0"]) and np.array_equal(tele.portal1, ref["portal1"])
    for k in ("x0", "x1", "kick", "delta0", "delta1"):
        g, r = getattr(tele, k), ref[k]
        assert np.array_equal(np.isnan(g), np.isnan(r)) and np.array_equal(g[~np.isnan(g)], r[~np.isnan(r)]), k


CASES = [dict(n=1, seed=1), dict(n=2, seed=2, side=3.0, axes="x"), dict(n=400, seed=1, side=10.0), dict(n=400, seed=2, side=10.0, lees_edwards=0.7, t=1.3, oblique=True),
         dict(n=600, seed=3, side=16.0, boxes=True, axes="x"), dict(n=400, seed=4, side=10.0, nfixed_frac=0.3), dict(n=500, seed=6, side=7.0, axes="y", lees_edwards=-0.9, t=4.0),
         dict(n=30000, seed=7, lees_edwards=0.4, t=0.9)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d-s%d" % (c["n"], c["seed"]))
def test_rb2d_portal_active_set_matches_oracle(gpu_ctx, oracle, case):
    s = scenes.rb2d_periodic(**case)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    dx_ref = o.update_portals(s["t"])
    assert np.array_equal(sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"]), dx_ref)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set_portals(s["q"], q1, "grid" if case["n"] > 3000 else "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    tele = sim.teleported()
    assert_equal(got, tele, ref, exact=not case.get("boxes", False))
    if case["n"] >= 400:
        assert tele.n_teleported > 5
    if case.get("lees_edwards", 0.0) != 0.0:
        assert np.any(got.type == 26)


def test_rb2d_portal_flow_resident_and_enforce(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb2d_periodic(5000, 13, lees_edwards=0.6, t=2.0, oblique=True)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    o.update_portals(s["t"])
    sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    q1, v1 = sb.VerletMap().flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = o.flow(1, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    ref = o.active_set_portals(s["q"], rq1)
    got = sim.computeActiveSet(s["q"], q1, resident=True)
    assert_equal(got, sim.teleported(), ref, exact=True)
    q = s["q"].copy()
    q.reshape(-1, 3)[:, :2] += np.random.default_rng(8).uniform(-0.45, 0.45, size=(5000, 2)) * s["side"]
    rq, rv = o.enforce_portals(q, s["v"])
    gq, gv = sim.enforcePeriodicBoundaryConditions(q, s["v"])
    assert np.array_equal(gq, rq) and np.array_equal(gv, rv) and np.any(gq != q) and np.any(gv != s["v"])


def test_rb2d_portal_unsupported_cases(gpu_ctx, oracle):
    import scisim_b200 as sb
    # boxes reaching the portals
    s = scenes.rb2d_periodic(300, 8, side=9.0, boxes=True)
    q = s["q"].reshape(-1, 3)
    q[:, :2] = np.random.default_rng(2).uniform(0.0, s["side"], size=q[:, :2].shape)
    s["q"] = q.ravel().copy()
    o = make_oracle(s)
    o.update_portals(0.0)
    assert not o.active_set_portals(s["q"], s["q"], "allpairs")["supported"]
    with pytest.raises(sb.SciSimB200Error):
        make_sim(s, gpu_ctx).computeActiveSet(s["q"], s["q"])
    # kinematic circles in teleported collisions
    s = scenes.rb2d_periodic(300, 9, side=9.0)
    s["fixed"][::3] = 1
    o = make_oracle(s)
    o.update_portals(0.0)
    assert not o.active_set_portals(s["q"], s["q"], "allpairs")["supported"]
    with pytest.raises(sb.SciSimB200Error):
        make_sim(s, gpu_ctx).computeActiveSet(s["q"], s["q"])


def test_rb2d_portals_cleared_restores_the_swept_path(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    sp = scenes.rb2d_periodic(300, 3, side=9.0)
    make_sim(sp, gpu_ctx).computeActiveSet(sp["q"], sp["q"])
    s = scenes.rb2d_random(1500, 4, kinds=("circle",))
    st = sb.RigidBody2DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_of_body"], s["fixed"], s["M"], s["g"], s["plane_x"], s["plane_n"])
    sim = sb.RigidBody2DSim(st, ctx=gpu_ctx)
    o = ob.RB2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    assert ref["supported"] and np.array_equal(got.candidates, ref["candidates"]) and np.array_equal(got.type, ref["type"])
    assert np.array_equal(got.i, ref["i"]) and np.array_equal(got.j, ref["j"]) and np.array_equal(got.n, ref["n"])
    with pytest.raises(sb.SciSimB200Error):
        sim.teleported()


def test_rb2d_portal_trajectory(gpu_ctx, oracle):
    """RigidBody2DSim::flow's portal bookkeeping over 12 steps without contact response (RigidBody2DSim.cpp:818-874)."""
    import scisim_b200 as sb
    s = scenes.rb2d_periodic(1500, 21, side=30.0, lees_edwards=2.0, vmax=20.0, dt=0.02)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    rq, rv = q.copy(), v.copy()
    crossed = 0
    for it in range(1, 13):
        assert np.array_equal(sim.updatePeriodicBoundaryConditionsStartOfStep(it, s["dt"]), o.update_portals(it * s["dt"]))
        q1, v1 = sb.SymplecticEulerMap().flow(q, v, sim, it, s["dt"])
        rq1, rv1 = o.flow(0, rq, rv, s["dt"])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
        ref = o.active_set_portals(rq, rq1)
        got = sim.computeActiveSet(q, q1, resident=True)
        assert_equal(got, sim.teleported(), ref, exact=True)
        q, v = sim.enforcePeriodicBoundaryConditions(q1, v1)
        rq, rv = o.enforce_portals(rq1, rv1)
        assert np.array_equal(q, rq) and np.array_equal(v, rv)
        crossed += int(np.any(q.reshape(-1, 3) != q1.reshape(-1, 3), axis=1).sum())
    assert crossed > 50
