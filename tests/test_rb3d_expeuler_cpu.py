"""CPU: the oracle's ExponentialEulerMap (rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.cpp:13-91).  The reference projects
the advanced orientation back to a rotation with U V^T from Eigen::JacobiSVD; the oracle computes that same (unique) orthogonal
polar factor with its own one-sided Jacobi.  Checked here against LAPACK's SVD (numpy) and for orthonormality; the explicit-Euler
parts are plain expressions checked against numpy bit for bit."""
import numpy as np

from scisim_b200 import scenes
from tests import oracle_binding as ob


def test_exponential_euler_matches_numpy(oracle):
    s = scenes.rb3d_random_boxes(400, 9, spin=True)
    n = 400
    o = ob.RB3DOracle(s)
    dt = 0.01
    q1, v1 = o.flow(4, s["q"], s["v"], dt)
    x0, R0 = s["q"][:3 * n].reshape(n, 3), s["q"][3 * n:].reshape(n, 3, 3)
    vl, w = s["v"][:3 * n].reshape(n, 3), s["v"][3 * n:].reshape(n, 3)
    assert np.array_equal(q1[:3 * n].reshape(n, 3), x0 + dt * vl)
    R1 = q1[3 * n:].reshape(n, 3, 3)
    worst = 0.0
    for b in range(n):
        A = np.stack([R0[b][:, j] + dt * np.cross(w[b], R0[b][:, j]) for j in range(3)], axis=1)
        U, S, Vt = np.linalg.svd(A)
        worst = max(worst, np.abs(U @ Vt - R1[b]).max())
        assert np.abs(R1[b] @ R1[b].T - np.eye(3)).max() < 1e-14
        assert abs(np.linalg.det(R1[b]) - 1.0) < 1e-13
    assert worst < 2e-14, worst   # LAPACK and the one-sided Jacobi agree to a few ulp
    m = s["m"]
    A = (1.0 / m)[:, None] * (m[:, None] * s["g"][None, :])
    assert np.array_equal(v1[:3 * n].reshape(n, 3), vl + dt * (0.0 + A))
    assert np.array_equal(v1[3 * n:].reshape(n, 3), w + dt * 0.0)
