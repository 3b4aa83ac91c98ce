"""Worker for tests/test_slab_multigpu.py (launched with torch.distributed.run, one process per GPU): several steps of a
RANDOMLY NUMBERED scene cut into x-quantile slabs (scisim_b200.slab.Ball2DSlabSim), the state advanced on the host between
steps so that a stale interval or halo cannot go unnoticed, one step uploaded mirrored in x so that every rank has to ask
for a re-partition; rank 0 merges the ranks' lists and compares them with the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import scisim_b200 as sb
    from scisim_b200.slab import Ball2DSlabSim, GpuSlabBackend
    from tests import slab_helpers as sh
    transport, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene = sh.random_numbered_scene(n, 11, box=max(4.0, np.sqrt(n) * 1.2))
    ctx = sb.Context(local)
    factory = lambda s, gids, lim, cap: GpuSlabBackend(ctx, s, 0, cap, gids=gids, x_limits=lim)
    sim = Ball2DSlabSim(scene, rank, world, dist, factory, transport=transport, ghost_cap=max(4096, n // 4))
    if rank == 0:
        from tests import oracle_binding as ob
        o = ob.Ball2DOracle(scene)
    q, v = scene["q"].copy(), scene["v"].copy()
    ok = True
    for step in range(steps):
        if step == 2:
            # mirror the scene: every body is now far from the slab that owns it
            x = q[0::2]
            q = q.copy()
            q[0::2] = x.min() + x.max() - x
        sim.upload(q, v)
        sim.step(0, scene["dt"])
        merged = sim.gather_merged(0)
        if rank == 0:
            q1, v1 = o.flow(0, q, v, scene["dt"])
            ref = o.active_set(q, q1, "grid")
            good = np.array_equal(merged["q1"], q1) and np.array_equal(merged["v1"], v1) and np.array_equal(merged["candidates"], ref["candidates"])
            for k in ("type", "i", "j", "n", "p"):
                good = good and np.array_equal(merged[k], ref[k])
            good = good and np.array_equal(merged["depth"], ref["depth"], equal_nan=True)
            if step >= 2:
                good = good and sim.n_partitions >= 2
            print("step %d: %s (candidates %d, active %d, partitions %d, transport %s)" % (step, "OK" if good else "MISMATCH", ref["candidates"].shape[0], ref["type"].shape[0], sim.n_partitions, sim.transport), flush=True)
            ok = ok and good
            nxt = [q1, v1]
        else:
            nxt = [None, None]
        dist.broadcast_object_list(nxt, src=0)
        q, v = nxt
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
