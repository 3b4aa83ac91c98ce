"""GPU, BASELINE.json full sizes, checked through size-independent properties (the oracle does not finish in seconds here):
  config 3  polydisperse 2-D ball gas, Verlet (4 M balls on one GPU here; 16 M is the 8-GPU total)
  config 4  3-D sphere box drop, 4 096 000 spheres, split_ham
  config 5  mixed sphere / box / mesh scene with 10 k bodies
Properties: ascending unique (i<j) candidate order; every active pair is a candidate; unit normals; depths <= 0;
detection idempotent; a random sample of bodies checked against brute force over ALL bodies, and (config 3) every pair inside a window of a few
thousand bodies against all-pairs in numpy (same predicate as the reference: closed AABB overlap, then the narrow-phase inequality)."""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu


def _check_order(cand):
    assert np.all(cand[:, 0] < cand[:, 1])
    key = cand[:, 0].astype(np.uint64) << np.uint64(32) | cand[:, 1].astype(np.uint64)
    assert np.all(key[1:] > key[:-1])
    return key


def test_config3_gas_4m(gpu_ctx):
    import scisim_b200 as sb
    n = 1 << 22
    s = scenes.ball2d_gas(n=n)
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"])
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.VerletMap(), s["dt"])
    q1, v1, a = sim.fetch()
    assert 1.0 * n < pc < 1.8 * n and 0.7 * n < a.n_body_body < 1.4 * n
    ck = _check_order(a.candidates)
    bb = a.type == 0
    ak = a.i[bb].astype(np.uint64) << np.uint64(32) | a.j[bb].astype(np.uint64)
    assert np.all(ak[1:] > ak[:-1]) and np.all(np.isin(ak, ck))
    assert np.all(np.abs(np.linalg.norm(a.n, axis=1) - 1.0) < 1e-12) and np.all(a.depth[~np.isnan(a.depth)] <= 0.0)
    # Verlet with zero gravity: q1 = q0 + dt*v0, v1 = v0 exactly
    assert np.array_equal(v1, s["v"] + 0.0) and np.array_equal(q1, s["q"] + s["dt"] * s["v"])
    # brute force for a sample of bodies against everything (swept AABB overlap == candidate)
    q0 = s["q"].reshape(-1, 2)
    q1r = q1.reshape(-1, 2)
    lo = np.minimum(q0, q1r) - s["r"][:, None]
    hi = np.maximum(q0, q1r) + s["r"][:, None]
    rng = np.random.default_rng(0)
    for i in rng.choice(n, size=128, replace=False):
        ov = np.all(~(hi[i] < lo) & ~(hi < lo[i]), axis=1)
        ov[i] = False
        mine = set(a.candidates[a.candidates[:, 0] == i, 1].tolist()) | set(a.candidates[a.candidates[:, 1] == i, 0].tolist())
        assert mine == set(np.nonzero(ov)[0].tolist())
    # every pair inside a window of the scene (a few thousand bodies): the exact candidate set among them, all-pairs in numpy, against the
    # GPU list restricted to pairs with both bodies in the window
    cx, cy = np.median(q0[:, 0]), np.median(q0[:, 1])
    half = 0.5 * np.sqrt(4000.0 / n) * (q0[:, 0].max() - q0[:, 0].min())
    inside = np.nonzero((np.abs(q0[:, 0] - cx) < half) & (np.abs(q0[:, 1] - cy) < half))[0]
    assert 2000 < inside.shape[0] < 8000
    L, H = lo[inside], hi[inside]
    ovm = np.all(~(H[:, None, :] < L[None, :, :]) & ~(H[None, :, :] < L[:, None, :]), axis=2)
    ii, jj = np.nonzero(np.triu(ovm, k=1))
    gi, gj = inside[ii], inside[jj]
    want = np.sort((np.minimum(gi, gj).astype(np.uint64) << np.uint64(32)) | np.maximum(gi, gj).astype(np.uint64))
    member = np.zeros(n, dtype=bool)
    member[inside] = True
    sel = member[a.candidates[:, 0]] & member[a.candidates[:, 1]]
    assert want.shape[0] > 1000 and np.array_equal(ck[sel], want)
    assert sim.step(sb.VerletMap(), s["dt"]) == (pc, pa)


def test_config4_spheres_4m(gpu_ctx):
    import scisim_b200 as sb
    s = scenes.rb3d_sphere_lattice(160, 160, 160)
    n = 160 ** 3
    st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"])
    sim = sb.RigidBody3DSim(st, ctx=gpu_ctx)
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.SplitHamMap(), s["dt"])
    q1, v1, a = sim.fetch()
    assert 12.5 * n < pc < 13.0 * n and 2.9 * n < a.n_body_body < 3.0 * n
    ck = _check_order(a.candidates)
    bb = a.type == 10
    ak = a.i[bb].astype(np.uint64) << np.uint64(32) | a.j[bb].astype(np.uint64)
    assert np.all(ak[1:] > ak[:-1]) and np.all(np.isin(ak, ck))
    assert np.all(np.abs(np.linalg.norm(a.n, axis=1) - 1.0) < 1e-12) and np.all(a.depth <= 0.0)
    assert a.n_plane == int((a.type == 14).sum()) and 5 * 160 * 160 * 0.4 < a.n_plane < 5 * 160 * 160 * 0.6   # jitter: about half of each boundary layer touches its plane
    # R unchanged (no spin), x1 = x0 + (dt*v0 + 0.5 dt^2 g) with v0 = 0
    assert np.array_equal(q1[3 * n:], s["q"][3 * n:])
    x0 = s["q"][:3 * n].reshape(-1, 3)
    x1 = q1[:3 * n].reshape(-1, 3)
    assert np.array_equal(x1[:, 0], x0[:, 0]) and np.all(x1[:, 1] < x0[:, 1])
    # sphere-sphere predicate for a sample of bodies against all
    r = 0.5
    rng = np.random.default_rng(1)
    for i in rng.choice(n, size=32, replace=False):
        d = x1 - x1[i]
        box = np.all(np.abs(d) <= 2 * r + 1e-9, axis=1)          # generous prefilter, exact test below
        idx = np.nonzero(box)[0]
        idx = idx[idx != i]
        lo_i, hi_i = x1[i] - r, x1[i] + r
        ov = np.all(~(hi_i < (x1[idx] - r)) & ~((x1[idx] + r) < lo_i), axis=1)
        mine = set(a.candidates[a.candidates[:, 0] == i, 1].tolist()) | set(a.candidates[a.candidates[:, 1] == i, 0].tolist())
        assert mine == set(idx[ov].tolist())


def test_config5_mixed_10k(gpu_ctx, oracle):
    """10 k bodies in three type-segregated blocks; small enough for the oracle (literal grid) to check everything."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    from tests.test_rb3d_gpu import assert_active_equal, make_sim
    s = scenes.rb3d_mixed_segregated(4700)
    n = s["geo_of_body"].shape[0]
    assert n > 9900
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    rq1, rv1 = o.flow(3, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], rq1, "grid")
    sim.upload(s["q"], s["v"])
    sim.step(sb.DMVMap(), s["dt"])
    q1, v1, a = sim.fetch()
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    assert_active_equal(a, ref)
