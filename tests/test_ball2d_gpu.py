"""GPU parity (pytest -m gpu): the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): candidate and active pair sets bit-exact in the reference's order; q1/v1
bit-exact; normals, points and depths within 1e-12 relative (they are in fact compared bit-exactly too).
"""
import json
import os

import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL = 1.0e-12


def make_sim(scene, ctx):
    import scisim_b200 as sb
    st = sb.Ball2DState(scene["r"], scene["m"], scene["g"], scene["plane_x"], scene["plane_n"], scene["drum_x"], scene["drum_r"])
    return sb.Ball2DSim(st, ctx=ctx)


def assert_active_equal(gpu, ref):
    assert gpu.n_candidates == ref["candidates"].shape[0]
    assert gpu.n_active == ref["type"].shape[0]
    if gpu.candidates is not None:
        assert np.array_equal(gpu.candidates, ref["candidates"])
    assert np.array_equal(gpu.type, ref["type"])
    assert np.array_equal(gpu.i, ref["i"])
    assert np.array_equal(gpu.j, ref["j"])
    for k in ("n", "p"):
        g, r = getattr(gpu, k), ref[k]
        assert np.all(np.abs(g - r) <= REL * np.maximum(1.0, np.abs(r)))
        assert np.array_equal(g, r), k + " within tolerance but not bit-identical"
    assert np.array_equal(np.isnan(gpu.depth), np.isnan(ref["depth"]))
    ok = ~np.isnan(ref["depth"])
    assert np.all(np.abs(gpu.depth[ok] - ref["depth"][ok]) <= REL * np.maximum(1.0, np.abs(ref["depth"][ok])))
    assert gpu.n_body_body == int((ref["type"] == 0).sum())
    assert gpu.n_drum == int((ref["type"] == 1).sum())
    assert gpu.n_plane == int((ref["type"] == 2).sum())


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (37, 3), (1000, 4), (20000, 5)])
def test_flow_bit_exact(gpu_ctx, oracle, n, seed, kind):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.ball2d_random(n, seed)
    sim = make_sim(s, gpu_ctx)
    umap = sb.SymplecticEulerMap() if kind == 0 else sb.VerletMap()
    q1, v1 = umap.flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = ob.Ball2DOracle(s).flow(kind, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (3, 9), (64, 3), (257, 6), (1000, 4), (5000, 5), (30000, 7)])
def test_active_set_matches_oracle(gpu_ctx, oracle, n, seed):
    from tests import oracle_binding as ob
    s = scenes.ball2d_random(n, seed, nplanes=3, ndrums=2)
    sim = make_sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid" if n > 3000 else "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    assert_active_equal(got, ref)


def test_active_set_static_state(gpu_ctx, oracle):
    """computeNumberOfCollisions calls computeActiveSet( q, q, v ) (ball2d/Ball2DSim.cpp:228): q0 == q1."""
    from tests import oracle_binding as ob
    s = scenes.ball2d_random(4000, 21)
    sim = make_sim(s, gpu_ctx)
    ref = ob.Ball2DOracle(s).active_set(s["q"], s["q"], "grid")
    assert_active_equal(sim.computeActiveSet(s["q"], s["q"]), ref)


def test_crowded_cell_fallback(gpu_ctx, oracle):
    """More partners per ball than the in-register list holds: the ordered-selection path must give the same list."""
    from tests import oracle_binding as ob
    rng = np.random.default_rng(8)
    n = 600
    s = scenes.ball2d_random(n, 8, nplanes=1, ndrums=0)
    s["q"] = rng.uniform(-0.5, 0.5, size=2 * n)      # everything on top of everything
    s["r"] = rng.uniform(0.2, 0.4, size=n)
    sim = make_sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], 1e-4)
    ref = o.active_set(s["q"], q1, "allpairs")
    assert ref["candidates"].shape[0] > 100000
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)


def test_polydisperse_gas(gpu_ctx, oracle):
    from tests import oracle_binding as ob
    s = scenes.ball2d_gas(n=40000)
    sim = make_sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, v1 = o.flow(1, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)


def test_lattice_config2_small(gpu_ctx, oracle):
    from tests import oracle_binding as ob
    s = scenes.ball2d_lattice(200, 150)
    sim = make_sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, v1 = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    got = sim.computeActiveSet(s["q"], q1)
    assert_active_equal(got, ref)
    n = 200 * 150
    assert 3.5 * n < got.n_candidates < 4.1 * n and 1.8 * n < got.n_body_body < 2.1 * n


def test_resident_step_equals_host_calls(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.ball2d_random(8000, 31, nplanes=2, ndrums=1)
    sim = make_sim(s, gpu_ctx)
    umap = sb.VerletMap()
    q1, v1 = umap.flow(s["q"], s["v"], sim, 1, s["dt"])
    a = sim.computeActiveSet(s["q"], q1)
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(umap, s["dt"])
    assert (pc, pa) == (a.n_candidates, a.n_active)
    rq1, rv1, b = sim.fetch()
    assert np.array_equal(rq1, q1) and np.array_equal(rv1, v1)
    for k in ("type", "i", "j", "n", "p", "candidates"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    assert np.array_equal(a.depth, b.depth, equal_nan=True)


def test_ccd_golden_cases_on_gpu(gpu_ctx):
    """The reference's 11 known-answer CCD cases (scisimtests/narrowphase_tests.cpp) through the CUDA narrow phase:
    two balls, q0/q1 as given; the pair is active iff the reference test expects a hit."""
    import scisim_b200 as sb
    cases = json.load(open(os.path.join(GOLD, "ccd_cases.json")))
    for c in cases:
        st = sb.Ball2DState([c["ra"], c["rb"]], [1.0, 1.0])
        sim = sb.Ball2DSim(st, ctx=gpu_ctx)
        q0 = np.array(c["q0a"] + c["q0b"])
        q1 = np.array(c["q1a"] + c["q1b"])
        a = sim.computeActiveSet(q0, q1)
        assert a.n_body_body == (1 if c["hit"] else 0), c["name"]
        if c["hit"]:
            assert a.n_candidates == 1 and a.i[0] == 0 and a.j[0] == 1


@pytest.mark.parametrize("name", ["spatial_grid_00", "spatial_grid_01", "spatial_grid_02"])
def test_aabb_fixtures_on_gpu(gpu_ctx, oracle, name):
    """ball2dtests/collision_detection_tests.cpp on the GPU: sg_candidate_pairs == all-pairs set."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    boxes = np.load(os.path.join(GOLD, "aabb_fixtures.npz"))[name]
    got = sb.SpatialGridDetector.getPotentialOverlaps(boxes, ctx=gpu_ctx)
    brute, _ = ob.aabb_overlaps(boxes, "allpairs")
    assert np.array_equal(got, brute)


def test_aabb_3d_random(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    rng = np.random.default_rng(5)
    lo = rng.uniform(0, 30, size=(20000, 3))
    boxes = np.hstack([lo, lo + rng.uniform(0.2, 2.5, size=(20000, 3))])
    got = sb.SpatialGridDetector.getPotentialOverlaps(boxes, ctx=gpu_ctx)
    ref, _ = ob.aabb_overlaps(boxes, "grid")
    assert np.array_equal(got, ref)


def test_empty_and_degenerate(gpu_ctx):
    import scisim_b200 as sb
    assert sb.SpatialGridDetector.getPotentialOverlaps(np.zeros((0, 4)), ctx=gpu_ctx).shape == (0, 2)
    st = sb.Ball2DState(np.zeros(0), np.zeros(0))
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    a = sim.computeActiveSet(np.zeros(0), np.zeros(0))
    assert a.n_active == 0 and a.n_candidates == 0
    # identical boxes: all pairs overlap
    boxes = np.tile(np.array([[0.0, 0.0, 1.0, 1.0]]), (40, 1))
    got = sb.SpatialGridDetector.getPotentialOverlaps(boxes, ctx=gpu_ctx)
    assert got.shape[0] == 40 * 39 // 2
    # touching boxes count as overlapping (strict '<' in AABB::overlaps)
    boxes = np.array([[0.0, 0.0, 1.0, 1.0], [1.0, 0.0, 2.0, 1.0], [2.0 + 1e-12, 0.0, 3.0, 1.0]])
    got = sb.SpatialGridDetector.getPotentialOverlaps(boxes, ctx=gpu_ctx)
    assert got.tolist() == [[0, 1]]


def test_million_ball_properties(gpu_ctx):
    """BASELINE config 2 at full size, checked through size-independent properties (no oracle at this size in
    the GPU suite): counts in the lattice's known range, ascending unique (i<j) order, every active pair is a
    candidate, normals unit length, depths <= 0, flow linear in dt, and idempotence of the detection."""
    import scisim_b200 as sb
    s = scenes.ball2d_lattice(1000, 1000)
    n = 1000 * 1000
    sim = make_sim(s, gpu_ctx)
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.SymplecticEulerMap(), s["dt"])
    q1, v1, a = sim.fetch()
    assert 3.9 * n < pc < 4.0 * n and 1.9 * n < a.n_body_body < 2.0 * n
    ck = a.candidates[:, 0].astype(np.uint64) << np.uint64(32) | a.candidates[:, 1].astype(np.uint64)
    assert np.all(a.candidates[:, 0] < a.candidates[:, 1]) and np.all(ck[1:] > ck[:-1])
    bb = a.type == 0
    ak = a.i[bb].astype(np.uint64) << np.uint64(32) | a.j[bb].astype(np.uint64)
    assert np.all(ak[1:] > ak[:-1]) and np.all(np.isin(ak, ck))
    assert np.all(np.abs(np.linalg.norm(a.n, axis=1) - 1.0) < 1e-12) and np.all(a.depth <= 0.0)
    # plane contacts: floor (plane 0) holds exactly the bottom row, walls the outer columns
    pl = a.type == 2
    assert a.n_plane == int(pl.sum()) and a.n_plane >= 1000
    assert np.all(np.diff(a.j[pl].astype(np.int64)) >= 0)
    # v1 = v0 + dt*g exactly for unit masses; q1 = q0 + dt*v1
    assert np.array_equal(v1.reshape(-1, 2)[:, 1], np.full(n, 0.0 + (s["dt"] * 1.0) * (0.0 + 1.0 * -9.81)))
    assert np.array_equal(q1, s["q"] + s["dt"] * v1)
    pc2, pa2 = sim.step(sb.SymplecticEulerMap(), s["dt"])
    assert (pc2, pa2) == (pc, pa)


@pytest.mark.parametrize("per_cell", [3.0, 7.0, 12.0])
def test_dense_scene_mask_regimes(gpu_ctx, oracle, per_cell):
    """Densities chosen so that a body's walk is ~27, ~63 and ~108 visits long: below, around and beyond what the 64-bit
    pass-1 masks cover, with candidate counts on both sides of the 8-key register sort and of the 12-entry local list."""
    from tests import oracle_binding as ob
    n = 5000
    s = scenes.ball2d_random(n, 31, nplanes=2, ndrums=1, rmin=0.2, rmax=0.4, vmax=2.0, dt=0.01)
    rng = np.random.default_rng(int(per_cell))
    half = 0.5 * np.sqrt(n * 0.85 * 0.85 / per_cell)   # cell width ~ largest swept extent ~ 0.85
    s["q"] = rng.uniform(-half, half, size=2 * n)
    sim = make_sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    assert ref["candidates"].shape[0] > 2 * n
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)


def test_active_set_on_resident_flow_result(gpu_ctx, oracle):
    """SG_IN_RESIDENT: the flow's (q0, q1) are reused on the device; same lists as with uploaded vectors, and an error
    when no flow preceded the call."""
    import scisim_b200 as sb
    s = scenes.ball2d_random(3000, 41, nplanes=2, ndrums=1)
    sim = make_sim(s, gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"], resident=True)
    q1, v1 = sim._flow(0, s["q"], s["v"], s["dt"], np.empty_like(s["q"]), np.empty_like(s["v"]))
    a = sim.computeActiveSet(s["q"], q1, resident=True)
    b = sim.computeActiveSet(s["q"], q1)
    assert a.n_active == b.n_active and a.n_candidates == b.n_candidates and a.n_active > 0
    for k in ("type", "i", "j", "n", "p", "candidates"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], q1, resident=True)   # the upload above replaced the resident pair


@pytest.mark.parametrize("scene_kind", ["lattice", "shuffled_lattice", "dense3", "dense7", "dense12"])
def test_dense_scenes_second_call_takes_the_record_staged_pass1(gpu_ctx, oracle, scene_kind):
    """A context whose last active set had >= 2.5 candidates per body runs the next pass 1 with the records staged by tensor copies
    (sg_bp_count_staged) instead of the float-box prefilter: same lists, first call and second, as the oracle -- on a lattice pile, on
    the same pile randomly numbered, and at densities on both sides of what the 64-bit masks and the sorting network hold."""
    from tests import oracle_binding as ob
    if scene_kind in ("lattice", "shuffled_lattice"):
        s = scenes.ball2d_lattice(180, 120)
        n = 180 * 120
        if scene_kind == "shuffled_lattice":
            perm = np.random.default_rng(3).permutation(n)
            s["q"] = np.ascontiguousarray(s["q"].reshape(n, 2)[perm].ravel()); s["v"] = np.ascontiguousarray(s["v"].reshape(n, 2)[perm].ravel())
            s["r"] = np.ascontiguousarray(s["r"][perm]); s["m"] = np.ascontiguousarray(s["m"][perm])
    else:
        per_cell = float(scene_kind[5:])
        n = 5000
        s = scenes.ball2d_random(n, 31, nplanes=2, ndrums=1, rmin=0.2, rmax=0.4, vmax=2.0, dt=0.01)
        rng = np.random.default_rng(int(per_cell))
        half = 0.5 * np.sqrt(n * 0.85 * 0.85 / per_cell)
        s["q"] = rng.uniform(-half, half, size=2 * n)
    sim = make_sim(s, gpu_ctx)
    o = ob.Ball2DOracle(s)
    q1, _ = o.flow(0, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    assert ref["candidates"].shape[0] >= 2.5 * n
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)   # float-box prefilter (nothing is known about the scene yet)
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)   # record-staged
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)
