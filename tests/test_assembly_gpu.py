"""GPU: device-side assembly for the solver's first step against the oracle, bit for bit: N = computeN with pruned zeros,
Q = N^T Minv N in Eigen's accumulation order and structure, the contact bases, and the constraint cache as a sorted-key join
(ImpactOperatorUtilities.cpp:10-48, ImpactMap.cpp:106-110, Ball2DSim.cpp:188-201, ConstraintCache.cpp:20-122)."""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu


def _sim(s, ctx):
    import scisim_b200 as sb
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
    return sb.Ball2DSim(st, ctx=ctx)


@pytest.mark.parametrize("scene_kind,n,seed", [("random", 50, 1), ("random", 4000, 2), ("lattice", 60, 3), ("asset", 0, 4)])
def test_assembly_equals_oracle(oracle, gpu_ctx, scene_kind, n, seed):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    if scene_kind == "random":
        s = scenes.ball2d_random(n, seed, nplanes=3, ndrums=1)
    elif scene_kind == "lattice":
        s = scenes.ball2d_lattice(n, n, seed=seed, with_planes=True)      # axis-aligned plane normals: exact zeros to prune
    else:
        s = scenes.ball2d_asset("pool_break_ten_deep")
        s["v"] = s["v"].copy(); s["v"][0] = 30.0                        # the cue ball reaches the rack in one step
    sim = _sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.SymplecticEulerMap(), s["dt"])
    q1, v1 = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    assert pa == ref["type"].shape[0] and pa > 0
    got, exp = sim.assemble(), o.assemble()
    assert exp["supported"] and got["n_constraints"] == pa and got["n_dofs"] == 2 * s["r"].shape[0]
    for k in ("n_outer", "n_inner", "n_values", "q_outer", "q_inner", "q_values", "bases"):
        assert np.array_equal(got[k], exp[k]), k
    if scene_kind == "lattice":
        assert got["n_outer"][-1] < 4 * int((ref["type"] == 0).sum()) + 2 * int((ref["type"] != 0).sum())   # something was pruned


def test_cache_join_equals_oracle(oracle, gpu_ctx):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.ball2d_random(6000, 9, nplanes=3, ndrums=1)
    sim = _sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    rng = np.random.default_rng(5)
    for step in range(3):
        sim.upload(q, v)
        pc, pa = sim.step(sb.SymplecticEulerMap(), s["dt"])
        q1, v1 = o.flow(0, q, v, s["dt"])
        ref = o.active_set(q, q1, "grid")
        assert pa == ref["type"].shape[0]
        got, hits = sim.getCachedConstraintImpulses(pa, 2)
        exp, ehits = o.cache_lookup(pa, 2)
        assert hits == ehits and np.array_equal(got, exp)
        assert (step == 0 and hits == 0) or (step > 0 and 0 < hits < pa)
        r = rng.normal(size=2 * pa)
        sim.cacheConstraints(r, 2)
        o.cache_store(r, 2)
        q, v = q1, v1
    sim.clearConstraintCache()
    got, hits = sim.getCachedConstraintImpulses(pa, 2)
    assert hits == 0 and not got.any()
