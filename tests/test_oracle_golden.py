"""CPU-only: pins the oracle (oracle/) against every stored-answer test the reference has for this path.

  * the 11 CCD cases of scisimtests/narrowphase_tests.cpp (tests/golden/ccd_cases.json)
  * the 3 AABB fixtures of ball2dtests/collision_detection_tests.cpp (tests/golden/aabb_fixtures.npz):
    as in the reference test, the literal spatial-grid algorithm must return exactly the all-pairs set
"""
import json
import os

import numpy as np
import pytest

from tests import oracle_binding as ob

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CCD_CASES = json.load(open(os.path.join(GOLD, "ccd_cases.json")))


@pytest.mark.parametrize("case", CCD_CASES, ids=[c["name"] for c in CCD_CASES])
def test_ccd_known_answers(case):
    c, hit, t = ob.ccd(case["q0a"], case["q1a"], case["ra"], case["q0b"], case["q1b"], case["rb"])
    if case["cexpected"] is not None:
        assert np.all(np.abs(c - np.array(case["cexpected"])) <= case["coeff_tol"])
    assert hit == case["hit"]
    if "toi" in case:
        if case["toi_tol"] == 0.0:
            assert t == case["toi"]
        else:
            assert abs(t - case["toi"]) <= case["toi_tol"]


@pytest.mark.parametrize("name", ["spatial_grid_00", "spatial_grid_01", "spatial_grid_02"])
def test_aabb_fixture_grid_equals_all_pairs(name):
    boxes = np.load(os.path.join(GOLD, "aabb_fixtures.npz"))[name]
    assert np.all(boxes[:, :2] < boxes[:, 2:])
    grid, _ = ob.aabb_overlaps(boxes, "grid")
    brute, _ = ob.aabb_overlaps(boxes, "allpairs")
    assert grid.shape[0] > 0
    assert np.array_equal(grid, brute)
    # ascending (i,j), i<j
    assert np.all(grid[:, 0] < grid[:, 1])
    key = grid[:, 0].astype(np.uint64) << np.uint64(32) | grid[:, 1].astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    # independent numpy brute force on the same predicate (closed intervals)
    n = boxes.shape[0]
    sel = np.random.default_rng(0).choice(n, size=400, replace=False)
    for i in sel[:50]:
        ov = np.all(~(boxes[i, 2:] < boxes[:, :2]) & ~(boxes[:, 2:] < boxes[i, :2]), axis=1)
        ov[i] = False
        mine = set(grid[grid[:, 0] == i, 1].tolist()) | set(grid[grid[:, 1] == i, 0].tolist())
        assert mine == set(np.nonzero(ov)[0].tolist())


def test_aabb_3d_grid_equals_all_pairs():
    rng = np.random.default_rng(5)
    lo = rng.uniform(0, 20, size=(3000, 3))
    boxes = np.hstack([lo, lo + rng.uniform(0.2, 2.5, size=(3000, 3))])
    grid, _ = ob.aabb_overlaps(boxes, "grid")
    brute, _ = ob.aabb_overlaps(boxes, "allpairs")
    assert grid.shape[0] > 1000
    assert np.array_equal(grid, brute)


def test_ball2d_flow_matches_formula():
    from scisim_b200 import scenes
    s = scenes.ball2d_random(500, seed=3)
    o = ob.Ball2DOracle(s)
    q0, v0, dt = s["q"], s["v"], s["dt"]
    m2 = np.repeat(s["m"], 2)
    g2 = np.tile(s["g"], 500)
    q1, v1 = o.flow(0, q0, v0, dt)
    F = 0.0 + m2 * g2
    v1e = v0 + (0.0 + (dt * (1.0 / m2)) * F)
    assert np.array_equal(v1, v1e) and np.array_equal(q1, q0 + dt * v1e)
    q1, v1 = o.flow(1, q0, v0, dt)
    s_ = (0.5 * dt) * (1.0 / m2)
    vh = v0 + (0.0 + s_ * F)
    assert np.array_equal(q1, q0 + dt * vh) and np.array_equal(v1, vh + s_ * F)


def test_ball2d_active_set_grid_equals_all_pairs_and_order():
    from scisim_b200 import scenes
    s = scenes.ball2d_random(1500, seed=11, nplanes=3, ndrums=2)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    a = o.active_set(s["q"], q1, "grid")
    b = o.active_set(s["q"], q1, "allpairs")
    for k in ("type", "i", "j", "n", "p", "candidates"):
        assert np.array_equal(a[k], b[k])
    assert np.array_equal(a["depth"], b["depth"], equal_nan=True)
    t = a["type"]
    assert np.all(np.diff(t.astype(np.int64)) >= 0)      # ball-ball | drums | planes
    assert (t == 0).sum() > 0 and (t == 1).sum() > 0 and (t == 2).sum() > 0
    assert a["candidates"].shape[0] >= (t == 0).sum()
    nn = np.linalg.norm(a["n"], axis=1)
    assert np.all(np.abs(nn - 1.0) < 1e-12)
    assert np.all(np.isnan(a["depth"][t == 1])) and np.all(a["depth"][t != 1] <= 0.0)


def test_rb3d_update_m_and_minv_and_the_two_mass_matrices(oracle):
    """oracle/rb3d.h: updateMandMinv (RigidBody3DState.cpp:428-462) against a plain numpy R diag R^T (rounding-level agreement, the
    column-major layout of the blocks exact), and the two matrices a flow can read -- as constructed (inertia block transposed,
    RigidBody3DState.cpp:165-182) or as updated: identical results without spin, last-bit differences with it."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(500, 11)
    o = ob.RB3DOracle(s)
    n = 500
    I, Ii = o.update_m_and_minv(s["q"])
    R = s["q"][3 * n:].reshape(n, 3, 3)
    I0 = np.asarray(s["I0"]).reshape(n, 3)
    ref = np.einsum("bik,bk,bjk->bij", R, I0, R)          # ref[b][r][c]
    refi = np.einsum("bik,bk,bjk->bij", R, 1.0 / I0, R)
    got = I.reshape(n, 3, 3).transpose(0, 2, 1)             # stored at 3 c + r
    goti = Ii.reshape(n, 3, 3).transpose(0, 2, 1)
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max() and np.abs(goti - refi).max() <= 1e-14 * np.abs(refi).max()
    assert np.any(got != got.transpose(0, 2, 1))            # symmetric only up to rounding: the layout matters
    for kind in (2, 3):
        a = o.flow(kind, s["q"], s["v"], s["dt"])
        b = o.flow(kind, s["q"], s["v"], s["dt"], m_updated=True)
        assert not np.array_equal(a[1], b[1]) and np.abs(a[1] - b[1]).max() <= 1e-13 * np.abs(a[1]).max()
        assert np.abs(a[0] - b[0]).max() <= 1e-13
    s0 = scenes.rb3d_random_spheres(300, 12)               # no spin: the angular momentum is zero either way
    o0 = ob.RB3DOracle(s0)
    a, b = o0.flow(2, s0["q"], s0["v"], s0["dt"]), o0.flow(2, s0["q"], s0["v"], s0["dt"], m_updated=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
