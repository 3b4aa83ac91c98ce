"""GPU (one device is enough): slab mode of the rigidbody3d sphere path -- BASELINE configs[3]'s decomposition -- with 2 and 3 slabs living
in one process on one GPU (mailboxes connected by address, phases in lockstep), randomly numbered spheres cut into x-quantile slabs,
several steps with the state advanced on the host; the merged lists must equal the single-scene oracle in the reference's order.  The real
one-process-per-GPU exchange (CUDA IPC) runs in tests/test_slab_multigpu.py."""
import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu


def _scene(n, seed):
    s = scenes.rb3d_random_spheres(n, seed, spin=True, nfixed_frac=0.0, nplanes=2)
    return s


@pytest.mark.parametrize("world,n,seed,kind,steps", [(2, 6000, 51, 2, 2), (3, 15000, 52, 3, 3)])
def test_rb3d_slabs_equal_oracle(oracle, world, n, seed, kind, steps):
    import scisim_b200 as sb
    from scisim_b200 import slab
    from tests import oracle_binding as ob
    s = _scene(n, seed)
    o = ob.RB3DOracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    rank_of, cuts, gids = slab.partition_quantiles_3d(q, world)
    ctxs = [sb.Context(0) for _ in range(world)]
    bks = [slab.RB3DSlabBackend(ctxs[r], s, gids[r], slab.slab_limits(cuts, r), ghost_cap=n) for r in range(world)]
    ptrs = [b.mailbox()[0] for b in bks]
    for r in range(world):
        for side, peer in ((0, r - 1), (1, r + 1)):
            if 0 <= peer < world:
                bks[r].connect(side, same_process_ptr=ptrs[peer])
    for step in range(steps):
        for r in range(world):
            bks[r].upload(*slab.rb3d_owned_state(q, v, gids[r]))
        for b in bks:
            b.flow(kind, s["dt"])
        for c in ctxs:
            c.synchronize()
        for b in bks:
            b.exchange(1)
        for c in ctxs:
            c.synchronize()
        for b in bks:
            b.exchange(2)
        parts, halo = [], 0
        for r in range(world):
            pc, pa = bks[r].detect()
            q1r, v1r, res = bks[r].fetch()
            assert (pc, pa) == (res["candidates"].shape[0], res["type"].shape[0])
            halo += sum(bks[r].ghosts)
            res["gids"], res["q1"], res["v1"] = gids[r], q1r, v1r
            parts.append(res)
        assert halo > 0
        rq1, rv1 = o.flow(kind, q, v, s["dt"])
        ref = o.active_set(q, rq1, "grid")
        assert ref["supported"] and ref["candidates"].shape[0] > 0 and (ref["type"] >= 14).any()
        q1, v1 = slab.rb3d_scatter_state(n, [(p["gids"], p["q1"], p["v1"]) for p in parts])
        tol = lambda a, b: np.all(np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.abs(b)))
        assert np.array_equal(q1[:3 * n], rq1[:3 * n]) and tol(q1, rq1) and tol(v1, rv1)
        merged = slab.merge_active_sets(parts, n_bodies=n)
        assert np.array_equal(merged["candidates"], ref["candidates"])
        for k in ("type", "i", "j", "aux"):
            assert np.array_equal(merged[k], ref[k]), k
        for k in ("n", "p"):
            assert tol(merged[k], ref[k]), k
        ok = ~np.isnan(ref["depth"])
        assert np.array_equal(np.isnan(merged["depth"]), ~ok) and tol(merged["depth"][ok], ref["depth"][ok])
        q, v = rq1, rv1
    for c in ctxs:
        c.close()
