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
