"""GPU parity (pytest -m gpu): SG_IN_RESIDENT and the resident step for rigidbody2d -- detection on the device copies the map left.

Added after the round's last full GPU run (the same flag is verified for ball2d and rigidbody3d in their own files); kept in a
file that sorts behind the executed ones so that `pytest -x` reaches every verified test first.
"""
import numpy as np
import pytest

from scisim_b200 import scenes
from tests.test_rb2d_gpu import make_sim

pytestmark = pytest.mark.gpu


def test_rb2d_active_set_on_resident_flow_result(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb2d_random(3000, 9, kinds=("circle", "box"))
    sim = make_sim(s, gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"], resident=True)
    q1, v1 = sim._flow(0, s["q"], s["v"], s["dt"])
    a = sim.computeActiveSet(s["q"], q1, resident=True)
    b = sim.computeActiveSet(s["q"], q1)
    assert a.n_active == b.n_active > 0 and a.n_candidates == b.n_candidates
    for k in ("type", "i", "j", "n", "p", "candidates"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_rb2d_resident_step_matches_oracle(gpu_ctx, oracle):
    """sg_rb2d_upload / step / fetch: the map and the detection on the device copies, against the oracle's flow + active set."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    for kind, umap in ((0, sb.SymplecticEulerMap()), (1, sb.VerletMap())):
        s = scenes.rb2d_random(4000, 17 + kind, kinds=("circle", "box"))
        sim = make_sim(s, gpu_ctx)
        o = ob.RB2DOracle(s)
        sim.upload(s["q"], s["v"])
        nc, na = sim.step(umap, s["dt"])
        q1, v1, got = sim.fetch()
        rq1, rv1 = o.flow(kind, s["q"], s["v"], s["dt"])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
        ref = o.active_set(s["q"], rq1, "grid")
        assert ref["supported"] and nc == ref["candidates"].shape[0] and na == ref["type"].shape[0] and na > 0
        assert np.array_equal(got.candidates, ref["candidates"])
        for k in ("type", "i", "j", "aux"):
            assert np.array_equal(getattr(got, k), ref[k]), k
        for k in ("n", "p", "depth"):  # rotated boxes go through sincos: 1e-12, as in tests/test_rb2d_gpu.py
            g, r = getattr(got, k), ref[k]
            ok = ~np.isnan(r)
            assert np.array_equal(np.isnan(g), np.isnan(r)) and np.all(np.abs(g[ok] - r[ok]) <= 1.0e-12 * np.maximum(1.0, np.abs(r[ok]))), k
        # the same step through host buffers gives the same lists
        h = sim.computeActiveSet(s["q"], q1)
        assert np.array_equal(h.i, got.i) and np.array_equal(h.j, got.j) and np.array_equal(h.n, got.n)
    with pytest.raises(sb.SciSimB200Error):
        st = make_sim(scenes.rb2d_random(10, 1), gpu_ctx)
        st.fetch()
