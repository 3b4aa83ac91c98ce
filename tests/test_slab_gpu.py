"""GPU (one device is enough): the slab-mode kernels (interval reduce, ordered halo pack / unpack, ghost-aware detection
with global indices) driven by hand for 2 and 3 slabs living on the same GPU, merged and compared with the single-scene
oracle.  The NCCL exchange itself is exercised by bench.py --gpus N and, for the host logic, by tests/test_slab_gloo.py."""
import numpy as np
import pytest

from tests import slab_helpers as sh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,n,seed", [(2, 4000, 3), (3, 9000, 4)])
def test_slab_kernels_merge_equals_oracle(oracle, world, n, seed):
    import torch
    import scisim_b200 as sb
    from scisim_b200.slab import REC_BYTES, GpuSlabBackend, merge_active_sets, partition_slab_major
    from tests import oracle_binding as ob
    scene = sh.slab_major_scene(n, seed)
    firsts, counts = partition_slab_major(n, world)
    ctxs = [sb.Context(0) for _ in range(world)]
    bks = [GpuSlabBackend(ctxs[r], sh.slab_of(scene, firsts[r], counts[r]), firsts[r], ghost_cap=n) for r in range(world)]
    ivs = [b.flow(0, scene["dt"]) for b in bks]
    for c in ctxs:
        c.synchronize()
    for r in range(world):
        for side, peer in ((0, r - 1), (1, r + 1)):
            if not (0 <= peer < world):
                continue
            buf = bks[r].pack(ivs[peer], side)
            ctxs[r].synchronize()
            dst = bks[peer].recv_buffer(1 - side)     # my side-1 list arrives as the peer's side-0 ghosts
            dst.copy_(buf)
            torch.cuda.synchronize()
            bks[peer].unpack(1 - side, dst)
        for peer in range(world):
            if abs(peer - r) > 1:
                assert bks[r].count_overlapping(ivs[peer]) == 0
    parts, halo = [], 0
    for r in range(world):
        pc, pa = bks[r].detect()
        q1, v1, res = bks[r].fetch()
        assert (pc, pa) == (res["candidates"].shape[0], res["type"].shape[0])
        halo += sum(bks[r].ghosts)
        res["q1"] = q1
        parts.append(res)
    assert halo > 0
    o = ob.Ball2DOracle(scene)
    q1, v1 = o.flow(0, scene["q"], scene["v"], scene["dt"])
    ref = o.active_set(scene["q"], q1, "grid")
    assert np.array_equal(np.concatenate([p["q1"] for p in parts]), q1)
    merged = merge_active_sets(parts, (scene["drum_x"].shape[0], scene["plane_x"].shape[0]))
    assert np.array_equal(merged["candidates"], ref["candidates"])
    for k in ("type", "i", "j", "n", "p"):
        assert np.array_equal(merged[k], ref[k]), k
    assert np.array_equal(merged["depth"], ref["depth"], equal_nan=True)
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("world,n,seed,steps", [(2, 4000, 5, 1), (3, 9000, 6, 3)])
def test_slab_peer_memory_exchange_equals_oracle(oracle, world, n, seed, steps):
    """The mailbox transport (sg_ball2d_slab_mailbox / connect / exchange): the slabs live in one process here, so the
    mailboxes are connected by address and the phases run in lockstep over the ranks; with one process per GPU the same
    calls go through CUDA IPC (bench.py --gpus N).  Several steps check the step-tagged flags."""
    import scisim_b200 as sb
    from scisim_b200.slab import GpuSlabBackend, merge_active_sets, partition_slab_major
    from tests import oracle_binding as ob
    scene = sh.slab_major_scene(n, seed)
    firsts, counts = partition_slab_major(n, world)
    ctxs = [sb.Context(0) for _ in range(world)]
    bks = [GpuSlabBackend(ctxs[r], sh.slab_of(scene, firsts[r], counts[r]), firsts[r], ghost_cap=n) for r in range(world)]
    ptrs = [b.mailbox()[0] for b in bks]
    for r in range(world):
        for side, peer in ((0, r - 1), (1, r + 1)):
            if 0 <= peer < world:
                bks[r].connect(side, same_process_ptr=ptrs[peer])
    o = ob.Ball2DOracle(scene)
    q, v = scene["q"].copy(), scene["v"].copy()
    for step in range(steps):
        for b in bks:
            b.flow(0, scene["dt"])
        for b in bks:
            b.exchange(1)
        for b in bks:
            b.exchange(2)
        parts, halo = [], 0
        for r in range(world):
            pc, pa = bks[r].detect()
            q1r, v1r, res = bks[r].fetch()
            halo += sum(bks[r].ghosts)
            res["q1"], res["v1"] = q1r, v1r
            parts.append(res)
        assert halo > 0
        q1, v1 = o.flow(0, q, v, scene["dt"])
        ref = o.active_set(q, q1, "grid")
        assert np.array_equal(np.concatenate([p["q1"] for p in parts]), q1)
        merged = merge_active_sets(parts, (scene["drum_x"].shape[0], scene["plane_x"].shape[0]))
        assert np.array_equal(merged["candidates"], ref["candidates"])
        for k in ("type", "i", "j", "n", "p"):
            assert np.array_equal(merged[k], ref[k]), k
        # next step starts from the unconstrained end state (no solver on this path)
        q, v = q1, v1
        vp = lambda a: a.ctypes.data_as(__import__("ctypes").c_void_p)
        for r in range(world):
            lo, hi = 2 * firsts[r], 2 * (firsts[r] + counts[r])
            qq, vv = np.ascontiguousarray(q[lo:hi]), np.ascontiguousarray(v[lo:hi])
            ctxs[r].check(ctxs[r].lib.sg_ball2d_upload(ctxs[r].h, vp(qq), vp(vv)))
    for c in ctxs:
        c.close()
