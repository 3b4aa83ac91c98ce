"""GPU parity for the rigidbody2d path (pytest -m gpu) vs oracle/rb2d.h: flows bit-exact; candidate / active sets
bit-exact in the reference's order; normals / points / depths within 1e-12 relative (rotated boxes go through sin/cos:
libm on the host, CUDA on the device)."""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
REL = 1.0e-12


def make_sim(s, ctx):
    import scisim_b200 as sb
    st = sb.RigidBody2DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_of_body"], s["fixed"], s["M"], s["g"], s["plane_x"], s["plane_n"])
    return sb.RigidBody2DSim(st, ctx=ctx)


def close(a, b):
    return np.all(np.abs(a - b) <= REL * np.maximum(1.0, np.abs(b)))


def assert_active_equal(gpu, ref, exact):
    assert ref["supported"]
    assert np.array_equal(gpu.candidates, ref["candidates"])
    assert gpu.n_active == ref["type"].shape[0]
    for k in ("type", "i", "j", "aux"):
        assert np.array_equal(getattr(gpu, k), ref[k]), k
    for k in ("n", "p"):
        assert close(getattr(gpu, k), ref[k]), k
        if exact:
            assert np.array_equal(getattr(gpu, k), ref[k]), k
    assert np.array_equal(np.isnan(gpu.depth), np.isnan(ref["depth"]))
    ok = ~np.isnan(ref["depth"])
    assert close(gpu.depth[ok], ref["depth"][ok])


@pytest.mark.parametrize("kind", [0, 1])
def test_rb2d_flow_bit_exact(gpu_ctx, oracle, kind):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb2d_random(5000, 3, kinds=("circle",))
    sim = make_sim(s, gpu_ctx)
    umap = sb.SymplecticEulerMap() if kind == 0 else sb.VerletMap()
    q1, v1 = umap.flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = ob.RB2DOracle(s).flow(kind, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    assert np.array_equal(q1.reshape(-1, 3)[s["fixed"] == 1], s["q"].reshape(-1, 3)[s["fixed"] == 1])


@pytest.mark.parametrize("kinds,n,seed,exact", [(("circle",), 3000, 1, True), (("box",), 2500, 2, False), (("circle", "box"), 4000, 3, False), (("circle", "box"), 2, 4, False)])
def test_rb2d_active_set(gpu_ctx, oracle, kinds, n, seed, exact):
    from tests import oracle_binding as ob
    s = scenes.rb2d_random(n, seed, kinds=kinds)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid" if n > 1000 else "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    assert_active_equal(got, ref, exact)
    if n >= 2500:
        assert len(np.unique(ref["type"])) >= 2


def test_rb2d_kinematic_box_is_an_error(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb2d_random(300, 5, kinds=("box",), box=2.0)
    s["fixed"][:] = 0
    s["fixed"][::7] = 1
    sim = make_sim(s, gpu_ctx)
    ref = ob.RB2DOracle(s).active_set(s["q"], s["q"], "allpairs")
    assert not ref["supported"]
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"])
