"""Worker for tests/test_slab_multigpu.py (torch.distributed.run, one process per GPU): an all-sphere rigidbody3d scene with random
numbering, cut into x-quantile slabs (scisim_b200.slab.RB3DSlabSim, CUDA IPC mailboxes), several steps with the state advanced on the
host, one step uploaded mirrored in x so that every rank has to ask for a re-partition; rank 0 merges and compares with the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import scisim_b200 as sb
    from scisim_b200 import scenes
    from scisim_b200.slab import RB3DSlabSim
    n, steps, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = scenes.rb3d_random_spheres(n, 61, spin=True, nfixed_frac=0.0, nplanes=2)
    ctx = sb.Context(local)
    sim = RB3DSlabSim(ctx, s, rank, world, dist, ghost_cap=max(8192, n // 2))
    if rank == 0:
        from tests import oracle_binding as ob
        o = ob.RB3DOracle(s)
    q, v = s["q"].copy(), s["v"].copy()
    ok = True
    tol = lambda a, b: np.all(np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.abs(b)))
    for step in range(steps):
        if step == 2:
            x = q[0:3 * n:3]
            q = q.copy()
            q[0:3 * n:3] = x.min() + x.max() - x
        sim.upload(q, v)
        sim.step(kind, s["dt"])
        m = sim.gather_merged(0)
        if rank == 0:
            rq1, rv1 = o.flow(kind, q, v, s["dt"])
            ref = o.active_set(q, rq1, "grid")
            good = tol(m["q1"], rq1) and tol(m["v1"], rv1) and np.array_equal(m["candidates"], ref["candidates"])
            for k in ("type", "i", "j", "aux"):
                good = good and np.array_equal(m[k], ref[k])
            good = good and tol(m["n"], ref["n"]) and tol(m["p"], ref["p"])
            if step >= 2:
                good = good and sim.n_partitions >= 2
            print("step %d: %s (candidates %d, active %d, partitions %d)" % (step, "OK" if good else "MISMATCH", ref["candidates"].shape[0], ref["type"].shape[0], sim.n_partitions), flush=True)
            ok = ok and good
            nxt = [rq1, rv1]
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
