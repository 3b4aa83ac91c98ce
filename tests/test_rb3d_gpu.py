"""GPU parity for the rigidbody3d path (pytest -m gpu): CUDA through the C ABI vs the CPU oracle (oracle/rb3d.h).

Bar: candidate / active sets bit-exact in the reference's order; DMV flow and every non-rotating flow bit-exact;
SplitHam with spin uses sin/cos (libm on the host, CUDA on the device) => 1e-12 relative; normals / points / depths
within 1e-12 relative (in practice bit-identical, which is also asserted)."""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
REL = 1.0e-12


def make_sim(s, ctx):
    import scisim_b200 as sb
    meshes = [sb.TriangleMesh(m["verts"], m["samples"], m["hull"], m["cell_delta"], m["dims"], m["origin"], m["sdf"]) for m in s["meshes"]]
    st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], meshes, s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"],
                             s.get("cyl_x"), s.get("cyl_axis"), s.get("cyl_r"))
    return sb.RigidBody3DSim(st, ctx=ctx)


def close(a, b):
    return np.all(np.abs(a - b) <= REL * np.maximum(1.0, np.abs(b)))


def assert_active_equal(gpu, ref, exact=True):
    assert ref["supported"]
    assert gpu.n_candidates == ref["candidates"].shape[0]
    if gpu.candidates is not None:
        assert np.array_equal(gpu.candidates, ref["candidates"])
    assert gpu.n_active == ref["type"].shape[0]
    for k in ("type", "i", "j", "aux"):
        assert np.array_equal(getattr(gpu, k), ref[k]), k
    for k in ("n", "p"):
        assert close(getattr(gpu, k), ref[k]), k
        if exact:
            assert np.array_equal(getattr(gpu, k), ref[k]), k + " within tolerance but not bit-identical"
    assert np.array_equal(np.isnan(gpu.depth), np.isnan(ref["depth"]))
    ok = ~np.isnan(ref["depth"])
    assert close(gpu.depth[ok], ref["depth"][ok])
    assert gpu.n_body_body == int((ref["type"] <= 13).sum())
    assert gpu.n_plane == int((ref["type"] >= 14).sum())


def kind_of(s):
    return 2 if s["map"] == "split_ham" else 3


@pytest.mark.parametrize("kind", [2, 3])
def test_flow_without_spin_bit_exact(gpu_ctx, oracle, kind):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_spheres(4000, 5, spin=False)
    sim = make_sim(s, gpu_ctx)
    umap = sb.SplitHamMap() if kind == 2 else sb.DMVMap()
    q1, v1 = umap.flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = ob.RB3DOracle(s).flow(kind, s["q"], s["v"], s["dt"])
    n = 4000
    assert np.array_equal(q1[:3 * n], rq1[:3 * n]) and np.array_equal(v1[:3 * n], rv1[:3 * n])
    if kind == 2:
        assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    else:
        assert close(q1, rq1) and close(v1, rv1)


def test_flow_dmv_with_spin(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(3000, 6, spin=True)
    sim = make_sim(s, gpu_ctx)
    q1, v1 = sb.DMVMap().flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = ob.RB3DOracle(s).flow(3, s["q"], s["v"], s["dt"])
    assert close(q1, rq1) and close(v1, rv1)
    # no transcendental functions on this path: expected to be bit-identical
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)


def test_flow_split_ham_with_spin(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(3000, 7, spin=True)
    sim = make_sim(s, gpu_ctx)
    q1, v1 = sb.SplitHamMap().flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = ob.RB3DOracle(s).flow(2, s["q"], s["v"], s["dt"])
    assert close(q1, rq1) and close(v1, rv1)
    n = 3000
    assert np.array_equal(q1[:3 * n], rq1[:3 * n]) and np.array_equal(v1[:3 * n], rv1[:3 * n])   # linear DoFs have no trig


@pytest.mark.parametrize("n,seed", [(1, 1), (2, 2), (50, 3), (3000, 4), (20000, 5)])
def test_spheres_active_set(gpu_ctx, oracle, n, seed):
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_spheres(n, seed, spin=True)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(kind_of(s), s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid" if n > 3000 else "allpairs")
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)


def test_sphere_lattice_config4_small(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_sphere_lattice(40, 30, 24)
    n = 40 * 30 * 24
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    rq1, rv1 = o.flow(2, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], rq1, "grid")
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.SplitHamMap(), s["dt"])
    q1, v1, a = sim.fetch()
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    assert_active_equal(a, ref)
    assert 11 * n < pc < 13 * n and 2.5 * n < a.n_body_body < 3.0 * n


@pytest.mark.parametrize("n,seed", [(2, 1), (40, 2), (1500, 3), (6000, 4)])
def test_boxes_active_set(gpu_ctx, oracle, n, seed):
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(n, seed)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(3, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    got = sim.computeActiveSet(s["q"], q1)
    assert_active_equal(got, ref)
    if n >= 1500:
        assert (ref["type"] == 12).sum() > 100 and (ref["type"] == 15).sum() > 100


def test_boxes_axis_aligned_stack(gpu_ctx, oracle):
    """Face-face contacts with exactly parallel axes (the 1.05 edge fudge and the > / >= tie-breaks decide here)."""
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(64, 9, spin=False, nfixed_frac=0.0, nplanes=1)
    n = 64
    ix, iy, iz = np.meshgrid(np.arange(4), np.arange(4), np.arange(4), indexing="ij")
    x = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1) * 0.95
    s["geo_half"][:] = 0.5
    s["q"][:3 * n] = x.ravel()
    s["q"][3 * n:] = np.tile(np.eye(3).ravel(), n)
    s["plane_x"][:] = [[0.0, -0.45, 0.0]]
    s["plane_n"][:] = [[0.0, 1.0, 0.0]]
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    ref = o.active_set(s["q"], s["q"], "allpairs")
    assert (ref["type"] == 12).sum() > 200
    assert_active_equal(sim.computeActiveSet(s["q"], s["q"]), ref)


@pytest.mark.parametrize("n,seed", [(2, 1), (30, 2), (150, 3)])
def test_meshes_active_set(gpu_ctx, oracle, n, seed):
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_meshes(n, seed)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(3, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "allpairs")
    got = sim.computeActiveSet(s["q"], q1)
    assert_active_equal(got, ref)
    if n >= 30:
        assert (ref["type"] == 12).sum() > 50 and (ref["type"] == 16).sum() > 10


def test_mixed_segregated_config5_small(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_mixed_segregated(300)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    rq1, rv1 = o.flow(3, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], rq1, "grid")
    sim.upload(s["q"], s["v"])
    pc, pa = sim.step(sb.DMVMap(), s["dt"])
    q1, v1, a = sim.fetch()
    assert np.array_equal(q1, rq1) and np.array_equal(v1, rv1)
    assert_active_equal(a, ref)
    assert set(np.unique(ref["type"]).tolist()) >= {10, 12}


def test_mixed_types_touching_is_an_error(gpu_ctx, oracle):
    """A sphere overlapping a box: the reference prints and exits (RigidBody3DSim.cpp:905-909); the ABI returns
    SG_ERR_UNSUPPORTED and the oracle reports the same."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_mixed_segregated(40)
    n = s["geo_of_body"].shape[0]
    box0 = int(np.nonzero(s["geo_type"][s["geo_of_body"]] == 0)[0][0])
    s["q"][3 * box0:3 * box0 + 3] = s["q"][0:3]        # move a box onto sphere 0
    sim = make_sim(s, gpu_ctx)
    ref = ob.RB3DOracle(s).active_set(s["q"], s["q"], "allpairs")
    assert not ref["supported"]
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"])


def test_empty_rb3d(gpu_ctx):
    import scisim_b200 as sb
    st = sb.RigidBody3DState([1], [0.5], [[0, 0, 0]], [0], [], np.zeros(0, np.uint32), np.zeros(0, np.uint8), np.zeros(0), np.zeros((0, 3)))
    sim = sb.RigidBody3DSim(st, ctx=gpu_ctx)
    a = sim.computeActiveSet(np.zeros(0), np.zeros(0))
    assert a.n_active == 0 and a.n_candidates == 0


def test_meshes_use_tma_bricks(gpu_ctx, oracle):
    """The mesh narrow phase must serve (most of) its distance-field lookups from TMA-staged bricks, and the
    result must be the oracle's either way."""
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_meshes(150, 3)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(3, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "allpairs")
    s0, d0 = sim.meshStats()
    got = sim.computeActiveSet(s["q"], q1)
    s1, d1 = sim.meshStats()
    assert s1 - s0 > 0, "no sweep was staged through TMA"
    assert (s1 - s0) > (d1 - d0), "most sweeps should fit the shared-memory brick: staged %d direct %d" % (s1 - s0, d1 - d0)
    assert_active_equal(got, ref)


@pytest.mark.parametrize("box", [3.0, 5.0])
def test_dense_spheres_sort_regimes(gpu_ctx, oracle, box):
    """Crowded all-sphere scenes: owned-candidate counts on both sides of the 16-key register sort, walks longer than the
    64-bit masks (the ordered-selection path) next to short ones."""
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_spheres(4000, 17, spin=False, box=box)
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(kind_of(s), s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    assert ref["candidates"].shape[0] > 8 * 4000
    assert_active_equal(sim.computeActiveSet(s["q"], q1), ref)


@pytest.mark.parametrize("kind", ["spheres", "meshes"])
def test_static_cylinders(gpu_ctx, oracle, kind):
    """Static cylinders (RigidBody3DSim.cpp:1504-1557): bodies live inside; spheres / mesh hull vertices that reach the
    wall are in contact; the contacts follow the plane contacts, cylinder-major.  Two cylinders with oblique,
    un-normalised axes."""
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_spheres(3000, 23, spin=False, nplanes=1) if kind == "spheres" else scenes.rb3d_random_meshes(60, 24, nplanes=1)
    x = s["q"][:3 * s["geo_of_body"].shape[0]].reshape(-1, 3)
    ext = float(np.abs(x).max())
    s["cyl_x"] = np.array([[0.1, -0.2, 0.3], [0.0, 0.0, 0.0]])
    s["cyl_axis"] = np.array([[0.2, 3.0, -0.1], [1.0, 0.1, 0.05]])
    s["cyl_r"] = np.array([0.8 * ext, 0.95 * ext])
    sim = make_sim(s, gpu_ctx)
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(kind_of(s), s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], q1, "grid")
    got = sim.computeActiveSet(s["q"], q1)
    code = 17 if kind == "spheres" else 18
    assert (ref["type"] == code).sum() > 20
    assert ref["type"][-1] == code and (ref["j"][ref["type"] == code] == 1).any() and (ref["j"][ref["type"] == code] == 0).any()
    assert_active_equal(got, ref)


def test_static_cylinder_with_free_box_is_an_error(gpu_ctx, oracle):
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(50, 25, nplanes=0)
    s["cyl_x"] = np.zeros((1, 3)); s["cyl_axis"] = np.array([[0.0, 1.0, 0.0]]); s["cyl_r"] = np.array([100.0])
    o = ob.RB3DOracle(s)
    q1, _ = o.flow(kind_of(s), s["q"], s["v"], s["dt"])
    assert not o.active_set(s["q"], q1, "grid")["supported"]
    sim = make_sim(s, gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], q1)


def test_rb3d_active_set_on_resident_flow_result(gpu_ctx, oracle):
    import scisim_b200 as sb
    s = scenes.rb3d_random_spheres(2000, 31, spin=True)
    sim = make_sim(s, gpu_ctx)
    with pytest.raises(sb.SciSimB200Error):
        sim.computeActiveSet(s["q"], s["q"], resident=True)
    q1, v1 = sim._flow(kind_of(s), s["q"], s["v"], s["dt"])
    a = sim.computeActiveSet(s["q"], q1, resident=True)
    b = sim.computeActiveSet(s["q"], q1)
    assert a.n_active == b.n_active > 0 and a.n_candidates == b.n_candidates
    for k in ("type", "i", "j", "n", "p", "candidates"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_flow_exponential_euler(gpu_ctx, oracle):
    """ExponentialEulerMap.cpp:13-91: explicit Euler parts bit-exact; the re-orthonormalised orientation (the reference: U V^T from
    Eigen::JacobiSVD; oracle: one-sided Jacobi; device: Newton's polar iteration -- three routes to the same unique factor) to 1e-12."""
    import scisim_b200 as sb
    from tests import oracle_binding as ob
    s = scenes.rb3d_random_boxes(3000, 8, spin=True)
    sim = make_sim(s, gpu_ctx)
    n = 3000
    q1, v1 = sb.ExponentialEulerMap().flow(s["q"], s["v"], sim, 1, s["dt"])
    rq1, rv1 = ob.RB3DOracle(s).flow(4, s["q"], s["v"], s["dt"])
    assert np.array_equal(q1[:3 * n], rq1[:3 * n]) and np.array_equal(v1, rv1)
    assert close(q1[3 * n:], rq1[3 * n:])
    R = q1[3 * n:].reshape(n, 3, 3)
    assert np.abs(np.einsum("bij,bkj->bik", R, R) - np.eye(3)).max() < 1e-14
    # and through the resident step
    sim.upload(s["q"], s["v"])
    sim.step(sb.ExponentialEulerMap(), s["dt"])
    q1s, v1s, _ = sim.fetch()
    assert np.array_equal(q1s, q1) and np.array_equal(v1s, v1)
