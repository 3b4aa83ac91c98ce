"""GPU parity (pytest -m gpu): ball2d with planar / Lees-Edwards portals through the C ABI vs the CPU oracle
(oracle/ball2d_portals.h, itself checked against the reference's PlanarPortal.cpp in tests/test_portals_cpu.py).

Bar: the extended candidate list (teleported boxes included), the teleported-box table, the active set in the reference's
order (regular | teleported | drums | planes) and the constructor arguments of the teleported contacts are bit-identical.
"""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu


def make_sim(scene, ctx):
    import scisim_b200 as sb
    pp = sb.PlanarPortal.from_arrays(scene["portals"])
    st = sb.Ball2DState(scene["r"], scene["m"], scene["g"], scene["plane_x"], scene["plane_n"], scene["drum_x"], scene["drum_r"], planar_portals=pp)
    return sb.Ball2DSim(st, ctx=ctx)


def make_oracle(scene):
    from tests import oracle_binding as ob
    o = ob.Ball2DOracle(scene)
    o.set_portals(scene["portals"])
    return o


def assert_portal_active_equal(gpu, tele, ref):
    assert gpu.n_candidates == ref["candidates"].shape[0]
    assert np.array_equal(gpu.candidates, ref["candidates"])
    assert np.array_equal(tele.box_body, ref["box_body"]) and np.array_equal(tele.box_portal, ref["box_portal"])
    assert gpu.n_active == ref["type"].shape[0]
    assert np.array_equal(gpu.type, ref["type"])
    assert np.array_equal(gpu.i, ref["i"]) and np.array_equal(gpu.j, ref["j"])
    assert np.array_equal(gpu.n, ref["n"]) and np.array_equal(gpu.p, ref["p"])
    assert np.array_equal(np.isnan(gpu.depth), np.isnan(ref["depth"]))
    ok = ~np.isnan(ref["depth"])
    assert np.array_equal(gpu.depth[ok], ref["depth"][ok])
    assert tele.n_regular == ref["n_regular"] and tele.n_teleported == ref["portal0"].shape[0]
    assert gpu.n_body_body == tele.n_regular + tele.n_teleported
    assert np.array_equal(tele.portal0, ref["portal0"]) and np.array_equal(tele.portal1, ref["portal1"])
    assert np.array_equal(tele.x0, ref["x0"]) and np.array_equal(tele.x1, ref["x1"]) and np.array_equal(tele.kick, ref["kick"])
    assert gpu.n_plane == int((ref["type"] == 2).sum()) and gpu.n_drum == int((ref["type"] == 1).sum())


CASES = [
    dict(n=1, seed=1, axes="xy"), dict(n=2, seed=2, axes="x", side=1.5),
    dict(n=300, seed=3, axes="xy", side=8.0), dict(n=300, seed=4, axes="x", side=8.0, oblique=True),
    dict(n=3000, seed=5, axes="xy", side=30.0, lees_edwards=0.75, t=3.7),
    dict(n=3000, seed=6, axes="y", side=30.0, lees_edwards=-1.25, t=11.3, oblique=True),
    dict(n=40000, seed=7, axes="xy", side=110.0, lees_edwards=0.4, t=0.9),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d-%s-le%g" % (c["n"], c["axes"], c.get("lees_edwards", 0.0)))
def test_portal_active_set_matches_oracle(gpu_ctx, oracle, case):
    kw = dict(case)
    s = scenes.ball2d_periodic(kw.pop("n"), kw.pop("seed"), **kw)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    dx_ref = o.update_portals(s["t"])
    dx = sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    assert np.array_equal(dx, dx_ref)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set_portals(s["q"], q1, "grid" if case["n"] > 3000 else "allpairs")
    assert ref is not None
    got = sim.computeActiveSet(s["q"], q1)
    tele = sim.teleported()
    assert_portal_active_equal(got, tele, ref)
    if case["n"] >= 300:
        assert tele.n_teleported > 0 and tele.box_body.shape[0] > 0
    if case.get("lees_edwards", 0.0) != 0.0:
        assert np.any(got.type == 4)


def test_portal_duplicates_and_corner_copies(gpu_ctx, oracle):
    """Dense small box: balls touching two portals at once (two teleported copies), pairs found through both of their
    copies (set semantics), pairs of two teleported copies that also collide un-teleported (skipped)."""
    s = scenes.ball2d_periodic(400, 11, side=6.0, rmin=0.2, rmax=0.45, axes="xy")
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    o.update_portals(0.0)
    sim.updatePeriodicBoundaryConditionsStartOfStep(0, 0.0)
    ref = o.active_set_portals(s["q"], s["q"], "allpairs")
    got = sim.computeActiveSet(s["q"], s["q"])
    tele = sim.teleported()
    assert_portal_active_equal(got, tele, ref)
    n = 400
    both = (ref["candidates"][:, 0] >= n) & (ref["candidates"][:, 1] >= n)
    assert both.sum() > 0
    bodies, counts = np.unique(tele.box_body, return_counts=True)
    assert np.any(counts == 2)


def test_portal_resident_step_and_fetch(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.ball2d_periodic(5000, 13, side=40.0, axes="xy", lees_edwards=0.6, t=2.0)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    o.update_portals(s["t"])
    sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    sim.upload(s["q"], s["v"])
    nc, na = sim.step(sb.VerletMap(), s["dt"])
    q1, v1, got = sim.fetch()
    rq1, rv1 = o.flow(1, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    ref = o.active_set_portals(s["q"], rq1)
    assert nc == ref["candidates"].shape[0] and na == ref["type"].shape[0]
    assert_portal_active_equal(got, sim.teleported(), ref)


def test_portal_enforce_matches_oracle(gpu_ctx, oracle):
    s = scenes.ball2d_periodic(20000, 15, lees_edwards=1.5, t=2.3, oblique=True)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    o.update_portals(s["t"])
    sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    rng = np.random.default_rng(8)
    q = s["q"] + rng.uniform(-0.45, 0.45, size=s["q"].shape) * s["side"]
    rq, rv = o.enforce_portals(q, s["v"])
    gq, gv = sim.enforcePeriodicBoundaryConditions(q, s["v"])
    assert np.array_equal(gq, rq) and np.array_equal(gv, rv)
    assert np.any(gq != q) and np.any(gv != s["v"])


def test_portal_both_planes_is_unsupported(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.ball2d_periodic(4, 1, side=1.0, rmin=0.6, rmax=0.7, axes="x")
    sim = make_sim(s, gpu_ctx)
    assert make_oracle(s).active_set_portals(s["q"], s["q"]) is None
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"])


def test_portals_cleared_restores_the_ccd_path(gpu_ctx, oracle):
    """A context that had portals goes back to swept boxes + CCD once they are removed."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    sp = scenes.ball2d_periodic(500, 3, side=10.0)
    make_sim(sp, gpu_ctx).computeActiveSet(sp["q"], sp["q"])
    s = scenes.ball2d_random(2000, 4, nplanes=3, ndrums=2)
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    assert np.array_equal(got.candidates, ref["candidates"]) and np.array_equal(got.type, ref["type"])
    assert np.array_equal(got.i, ref["i"]) and np.array_equal(got.j, ref["j"]) and np.array_equal(got.n, ref["n"])
    with pytest.raises(sb.SciSimB200Error):
        sim.teleported()

