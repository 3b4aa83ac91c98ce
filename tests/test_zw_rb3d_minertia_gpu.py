"""GPU parity (pytest -m gpu): RigidBody3DState::updateMandMinv on the device and the flows that read the matrix it leaves
(SG_MAP_M_UPDATED, include/scisim_b200.h) against the oracle (oracle/rb3d.h: updateMandMinv, flow( ..., m_updated )).

Written after the round's GPU minutes were spent (tests/test_gpu_tests_replay.py replays this file's bodies on the CPU); it sorts
behind the files the last full GPU run executed.

Bar: the 18 N block values bit-identical; DMV flows bit-identical; SplitHam flows with spin within 1e-12 (sin / cos), as in
tests/test_rb3d_gpu.py.
"""
import numpy as np
import pytest

from scisim_b200 import scenes
from tests.test_rb3d_gpu import make_sim

pytestmark = pytest.mark.gpu
REL = 1.0e-12


def make_oracle(s):
    from tests import oracle_binding as ob
    return ob.RB3DOracle(s)


def close(a, b):
    return np.all(np.abs(a - b) <= REL * np.maximum(1.0, np.abs(b)))


@pytest.mark.parametrize("n,seed", [(1, 1), (777, 2), (20000, 3)])
def test_update_m_and_minv_matches_oracle(gpu_ctx, oracle, n, seed):
    s = scenes.rb3d_random_boxes(n, seed)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    I, Ii = sim.updateMandMinv(s["q"])
    rI, rIi = o.update_m_and_minv(s["q"])
    assert np.array_equal(I, rI) and np.array_equal(Ii, rIi)
    # the blocks are the world-space inertia and its inverse
    B, Bi = I.reshape(-1, 3, 3), Ii.reshape(-1, 3, 3)
    assert np.abs(np.einsum("bij,bjk->bik", B, Bi) - np.eye(3)).max() < 1.0e-9


def test_flows_after_the_first_read_the_updated_matrix(gpu_ctx, oracle):
    """RigidBody3DSim::flow over 6 steps without contact response (RigidBody3DSim.cpp:430-443): map, then updateMandMinv; the first
    flow multiplies v0 by M as constructed, the others by M as updated -- the two differ in the last bit for spinning boxes."""
    import scisim_b200 as sb
    s = scenes.rb3d_random_boxes(3000, 5)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    a = o.flow(3, s["q"], s["v"], s["dt"], m_updated=False)
    b = o.flow(3, s["q"], s["v"], s["dt"], m_updated=True)
    assert not np.array_equal(a[1], b[1])  # the scene tells the two matrices apart
    q, v = s["q"].copy(), s["v"].copy()
    rq, rv = q.copy(), v.copy()
    for it in range(1, 7):
        q1, v1 = sb.DMVMap().flow(q, v, sim, it, s["dt"])
        rq1, rv1 = o.flow(3, rq, rv, s["dt"], m_updated=it > 1)
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1), it
        I, Ii = sim.updateMandMinv()  # the device copy of q1
        rI, rIi = o.update_m_and_minv(rq1)
        assert np.array_equal(I, rI) and np.array_equal(Ii, rIi), it
        q, v, rq, rv = q1, v1, rq1, rv1


def test_split_ham_with_the_updated_matrix(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb3d_random_boxes(2000, 6)
    sim = make_sim(s, gpu_ctx)
    o = make_oracle(s)
    sim.updateMandMinv(s["q"])
    q1, v1 = sb.SplitHamMap().flow(s["q"], s["v"], sim, 2, s["dt"])
    rq1, rv1 = o.flow(2, s["q"], s["v"], s["dt"], m_updated=True)
    assert close(q1, rq1) and close(v1, rv1)
    # resident stepping carries the flag as well
    sim.upload(s["q"], s["v"])
    sim.step(sb.DMVMap(), s["dt"])
    q1, v1, _ = sim.fetch()
    rq1, rv1 = o.flow(3, s["q"], s["v"], s["dt"], m_updated=True)
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)


def test_update_without_a_configuration_is_an_error(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb3d_random_boxes(50, 7)
    sim = make_sim(s, gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.updateMandMinv()
