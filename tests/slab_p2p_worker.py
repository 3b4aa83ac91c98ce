"""Worker for tests/test_slab_multigpu.py (launched with torch.distributed.run, one process per GPU): several steps of a
slab-major scene through the peer-memory transport, the state advanced on the host between steps so that a stale
interval or halo cannot go unnoticed; rank 0 merges the ranks' lists and compares them with the oracle."""
import os
import pickle
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import scisim_b200 as sb
    from scisim_b200.slab import Ball2DSlabs, GpuSlabBackend, merge_active_sets, partition_slab_major
    from tests import slab_helpers as sh
    transport, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = sh.slab_major_scene(n, 11)
    firsts, counts = partition_slab_major(n, world)
    ctx = sb.Context(local)
    backend = GpuSlabBackend(ctx, sh.slab_of(scene, firsts[rank], counts[rank]), firsts[rank], ghost_cap=n)
    slabs = Ball2DSlabs(backend, rank, world, dist, transport=transport)
    if rank == 0:
        from tests import oracle_binding as ob
        o = ob.Ball2DOracle(scene)
    q, v = scene["q"].copy(), scene["v"].copy()
    ok = True
    for step in range(steps):
        slabs.step(0, scene["dt"])
        q1r, v1r, res = backend.fetch()
        res["q1"], res["v1"] = q1r, v1r
        blob = pickle.dumps(res)
        gathered = [None] * world
        dist.all_gather_object(gathered, blob)
        if rank == 0:
            parts = [pickle.loads(b) for b in gathered]
            q1, v1 = o.flow(0, q, v, scene["dt"])
            ref = o.active_set(q, q1, "grid")
            merged = merge_active_sets(parts, (scene["drum_x"].shape[0], scene["plane_x"].shape[0]))
            good = np.array_equal(np.concatenate([p["q1"] for p in parts]), q1) and np.array_equal(merged["candidates"], ref["candidates"])
            for k in ("type", "i", "j", "n", "p"):
                good = good and np.array_equal(merged[k], ref[k])
            halo = sum(int(p["candidates"].shape[0]) for p in parts)
            print("step %d: %s (candidates %d, active %d)" % (step, "OK" if good else "MISMATCH", ref["candidates"].shape[0], ref["type"].shape[0]), flush=True)
            ok = ok and good
            nxt = [q1, v1]
        else:
            nxt = [None, None]
        dist.broadcast_object_list(nxt, src=0)
        q, v = nxt
        lo, hi = 2 * firsts[rank], 2 * (firsts[rank] + counts[rank])
        qq, vv = np.ascontiguousarray(q[lo:hi]), np.ascontiguousarray(v[lo:hi])
        import ctypes as C
        ctx.check(ctx.lib.sg_ball2d_upload(ctx.h, qq.ctypes.data_as(C.c_void_p), vv.ctypes.data_as(C.c_void_p)))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
