"""GPU parity (pytest -m gpu): a ball2d portal trajectory -- only calls that tests/test_portals_gpu.py already verified on the
B200 (flow, portal active set, enforce), composed over many steps.  Added after the round's GPU minutes were spent, so it sorts
behind every file that has been executed on the GPU (and before the rigidbody2d / rigidbody3d portal files)."""
import numpy as np
import pytest

from scisim_b200 import scenes
from tests.test_portals_gpu import assert_portal_active_equal, make_oracle, make_sim

pytestmark = pytest.mark.gpu


def test_portal_trajectory_many_steps(gpu_ctx, oracle):
    """Ball2DSim::flow's portal bookkeeping over 25 steps without contact response (Ball2DSim.cpp:307-325): move the
    Lees-Edwards portals to t, integrate, detect, teleport the balls that left (with the velocity kick).  Balls cross the
    boundaries repeatedly; state, portal offsets and active sets stay bit-identical to the oracle's at every step."""
    import scisim_b200 as sb
    s = scenes.ball2d_periodic(3000, 31, side=30.0, axes="xy", lees_edwards=2.5, vmax=25.0, dt=0.02)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    rq, rv = q.copy(), v.copy()
    crossed = 0
    for it in range(1, 26):
        dx = sim.updatePeriodicBoundaryConditionsStartOfStep(it, s["dt"])
        assert np.array_equal(dx, o.update_portals(it * s["dt"]))
        q1, v1 = sb.SymplecticEulerMap().flow(q, v, sim, it, s["dt"])
        rq1, rv1 = o.flow(0, rq, rv, s["dt"])
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
        ref = o.active_set_portals(rq, rq1)
        assert ref is not None
        got = sim.computeActiveSet(q, q1, resident=True)
        assert_portal_active_equal(got, sim.teleported(), ref)
        q, v = sim.enforcePeriodicBoundaryConditions(q1, v1)
        rq, rv = o.enforce_portals(rq1, rv1)
        assert np.array_equal(q, rq) and np.array_equal(v, rv)
        crossed += int(np.any(q.reshape(-1, 2) != q1.reshape(-1, 2), axis=1).sum())
    assert crossed > 200
