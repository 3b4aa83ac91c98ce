"""GPU: BASELINE configs[0] -- the reference's own bundled ball2d scenes -- 100 steps of flow + active set through the C ABI
against the oracle, bit for bit, the state advanced by the unconstrained map (there is no contact response on this path, so
the balls interpenetrate more and more: contact counts grow step by step, which is the point).  Integrator forced to
symplectic_euler, dt from the scene file; the 6 079-ball scene also runs with its planar portal."""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu


def _check(a, ref):
    assert np.array_equal(a.candidates, ref["candidates"])
    for k in ("type", "i", "j", "n", "p"):
        assert np.array_equal(getattr(a, k), ref[k]), k
    assert np.array_equal(a.depth, ref["depth"], equal_nan=True)


@pytest.mark.parametrize("name,umap", [("pool_break_ten_deep", "symplectic_euler"), ("different_friction", "symplectic_euler"), ("different_friction", "verlet")])
def test_bundled_scene_100_steps(oracle, gpu_ctx, name, umap):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.ball2d_asset(name, map=umap)
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    m = sb.SymplecticEulerMap() if umap == "symplectic_euler" else sb.VerletMap()
    o = ob.Ball2DOracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    total = 0
    for it in range(100):
        q1, v1 = m.flow(q, v, sim, it + 1, s["dt"])
        rq1, rv1 = o.flow(m.kind, q, v, s["dt"])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1), "flow, step %d" % it
        a = sim.computeActiveSet(q, q1, resident=(it % 2 == 0))
        ref = o.active_set(q, rq1, "grid")
        _check(a, ref)
        total += a.n_active
        q, v = q1, v1
    assert total > 0


def test_bundled_scene_with_its_planar_portal(oracle, gpu_ctx):
    """different_friction.xml as it is written: planes 1 and 2 are a planar portal (periodic in x), plane 0 is the floor."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.ball2d_asset("different_friction", portal=True)
    assert s["plane_x"].shape[0] == 1 and s["portals"]["v"].shape[0] == 1
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"], planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    o = ob.Ball2DOracle(s)
    o.set_portals(s["portals"])
    m = sb.SymplecticEulerMap()
    q, v = s["q"].copy(), s["v"].copy()
    tele = 0
    for it in range(60):
        sim.updatePeriodicBoundaryConditionsStartOfStep(it + 1, s["dt"])
        o.update_portals((it + 1) * s["dt"])
        q1, v1 = m.flow(q, v, sim, it + 1, s["dt"])
        rq1, rv1 = o.flow(0, q, v, s["dt"])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
        a = sim.computeActiveSet(q, q1)
        ref = o.active_set_portals(q, rq1, "grid")
        assert ref is not None
        _check(a, ref)
        tele += int((a.type >= 3).sum())
        q, v = sim.enforcePeriodicBoundaryConditions(q1, v1)
        rq, rv = o.enforce_portals(rq1, rv1)
        assert np.array_equal(q, rq) and np.array_equal(v, rv)
    sim.state = None
