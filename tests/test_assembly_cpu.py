"""CPU: the oracle's N, Q = N^T Minv N, contact bases and constraint cache (oracle/assembly2d.h) against plain dense linear algebra: Q's pattern
and values must be those of the dense product (values to rounding: the sparse product adds the same terms, in Eigen's order), N must be the
pruned gradient matrix, the cache must behave like the three std::map of ball2d/ConstraintCache.cpp."""
import numpy as np

from scisim_b200 import scenes
from tests import oracle_binding as ob


def test_oracle_assembly_against_dense(oracle):
    s = scenes.ball2d_random(300, 17, nplanes=3, ndrums=1)
    o = ob.Ball2DOracle(s)
    q1, v1 = o.flow(0, s["q"], s["v"], s["dt"])
    a = o.active_set(s["q"], q1, "allpairs")
    na, n = a["type"].shape[0], 300
    assert na > 20 and (a["type"] == 0).any() and (a["type"] == 2).any()
    asm = o.assemble()
    assert asm["supported"]
    N = np.zeros((2 * n, na))
    for c in range(na):
        i, j, t, (nx, ny) = int(a["i"][c]), int(a["j"][c]), int(a["type"][c]), a["n"][c]
        N[2 * i, c], N[2 * i + 1, c] = nx, ny
        if t == 0:
            N[2 * j, c], N[2 * j + 1, c] = -nx, -ny
    Nd = np.zeros_like(N)
    for c in range(na):
        rows = asm["n_inner"][asm["n_outer"][c]:asm["n_outer"][c + 1]]
        assert np.all(np.diff(rows) > 0)
        Nd[rows, c] = asm["n_values"][asm["n_outer"][c]:asm["n_outer"][c + 1]]
    assert np.array_equal(Nd, N) and np.all(asm["n_values"] != 0.0)
    minv = np.repeat(1.0 / s["m"], 2)
    Q = N.T @ (minv[:, None] * N)
    Qd = np.zeros_like(Q)
    pattern = np.zeros_like(Q, dtype=bool)
    for d in range(na):
        rows = asm["q_inner"][asm["q_outer"][d]:asm["q_outer"][d + 1]]
        assert np.all(np.diff(rows) > 0)
        Qd[rows, d] = asm["q_values"][asm["q_outer"][d]:asm["q_outer"][d + 1]]
        pattern[rows, d] = True
    assert np.abs(Qd - Q).max() <= 1e-14 * np.abs(Q).max()   # same terms, possibly another order: differences of rounding size (entries may nearly cancel)
    structural = (np.abs(N.T) > 0).astype(float) @ (np.abs(N) > 0).astype(float) > 0
    assert np.array_equal(pattern, structural)
    B = asm["bases"].reshape(na, 4)
    assert np.array_equal(B[:, :2], a["n"]) and np.array_equal(B[:, 2], -a["n"][:, 1]) and np.array_equal(B[:, 3], a["n"][:, 0])


def test_oracle_cache_join(oracle):
    s = scenes.ball2d_random(400, 18, nplanes=2, ndrums=1)
    o = ob.Ball2DOracle(s)
    q1, v1 = o.flow(0, s["q"], s["v"], s["dt"])
    a = o.active_set(s["q"], q1, "allpairs")
    na = a["type"].shape[0]
    r = np.arange(1, 2 * na + 1, dtype=np.float64)
    o.cache_store(r, 2)
    got, hits = o.cache_lookup(na, 2)
    assert hits == na and np.array_equal(got, r)
    # next step: some contacts persist, some are new
    q2, v2 = o.flow(0, q1, v1, s["dt"])
    b = o.active_set(q1, q2, "allpairs")
    nb = b["type"].shape[0]
    got, hits = o.cache_lookup(nb, 2)
    key = lambda t, i, j: (int(t), int(i), int(j))
    old = {key(a["type"][c], a["i"][c], a["j"][c]): r[2 * c:2 * c + 2] for c in range(na)}
    exp = np.concatenate([old.get(key(b["type"][c], b["i"][c], b["j"][c]), np.zeros(2)) for c in range(nb)]) if nb else np.zeros(0)
    assert np.array_equal(got, exp) and 0 < hits < nb
