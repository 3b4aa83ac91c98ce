"""BASELINE configs[3] -- rigidbody3d 160^3 = 4 096 000-sphere box drop, static planes, split_ham ("Verlet") -- on 1/2/4/8 GPUs:
  python profiles/config4_multi.py                                            (1 GPU)
  python -m torch.distributed.run --nproc-per-node N ... profiles/config4_multi.py   (N x-quantile slabs, one process per GPU)
Resident-step timing like bench.py (CUDA events on the library stream, L2 flushed per step, max over ranks) and a parity check of a reduced
lattice through the same path against the CPU oracle.  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=160)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--map", default="split_ham")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    saved = os.dup(1); os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import scisim_b200 as sb
    from scisim_b200 import scenes
    from scisim_b200.slab import RB3DSlabSim
    ctx = sb.Context(local)
    kind = 2 if args.map == "split_ham" else 3
    umap = sb.SplitHamMap() if kind == 2 else sb.DMVMap()

    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    def run(scene, check):
        if world == 1:
            from tests.test_rb3d_gpu import make_sim
            sim = make_sim(scene, ctx)
            sim.upload(scene["q"], scene["v"])
            step = lambda: sim.step(umap, scene["dt"])
        else:
            sim = RB3DSlabSim(ctx, scene, rank, world, dist)
            sim.upload(scene["q"], scene["v"])
            step = lambda: sim.step(kind, scene["dt"])
        if check:
            step()
            if world == 1:
                q1, v1, a = sim.fetch()
                got = {"q1": q1, "v1": v1, "candidates": a.candidates, "type": a.type, "i": a.i, "j": a.j, "aux": a.aux, "n": a.n, "p": a.p}
            else:
                got = sim.gather_merged(0)
                sim.backend.disconnect()
            if rank != 0:
                return None
            from tests import oracle_binding as ob
            o = ob.RB3DOracle(scene)
            rq1, rv1 = o.flow(kind, scene["q"], scene["v"], scene["dt"])
            ref = o.active_set(scene["q"], rq1, "grid")
            tol = lambda a, b: bool(np.all(np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.abs(b))))
            bad = [k for k in ("candidates", "type", "i", "j", "aux") if not np.array_equal(got[k], ref[k])]
            bad += [k for k, r in (("q1", rq1), ("v1", rv1), ("n", ref["n"]), ("p", ref["p"])) if not tol(got[k], r)]
            return {"ok": not bad, "mismatch": bad, "bodies": int(scene["m"].shape[0]), "candidates": int(ref["candidates"].shape[0]), "active": int(ref["type"].shape[0])}
        barrier()
        for _ in range(args.warmup):
            ctx.flush_l2(); pc, pa = step()
        barrier()
        ms = []
        for _ in range(args.steps):
            ctx.flush_l2(); ctx.timer_begin(); pc, pa = step(); ms.append(ctx.timer_end())
        barrier()
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(args.steps):
            ctx.flush_l2(); step()
        prof = ctx.profile(); ctx.profile_enable(False)
        return pc, pa, ms, prof

    parity = run(scenes.rb3d_sphere_lattice(30, 30, 30), True)
    barrier()
    scene = scenes.rb3d_sphere_lattice(args.side, args.side, args.side)
    pc, pa, ms, prof = run(scene, False)
    t = sum(ms) / 1e3
    if world > 1:
        tt = torch.tensor([t], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t = float(tt[0])
        ww = torch.tensor([float(pc), float(pa)], dtype=torch.float64, device="cuda"); dist.all_reduce(ww, op=dist.ReduceOp.SUM); pc, pa = float(ww[0]), float(ww[1])
    if rank == 0:
        line = {"workload": "configs[3]: rigidbody3d %d-sphere box drop (lattice spacing 0.99, r = 0.5), floor + 4 walls, %s" % (args.side ** 3, args.map), "n_gpus": world, "scaling": "strong",
                "value": (pc + pa) * args.steps / t, "unit": "pairs/s", "ms_per_step": 1e3 * t / args.steps, "steps": args.steps, "warmup": args.warmup, "candidates": pc, "active": pa,
                "parity_check": parity, "kernels_us_rank0": {k: round(1e3 * v[1] / args.steps, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                "kernels_alg_GBps_rank0": {k: round(v[2] / (v[1] * 1e-3) / 1e9) for k, v in prof.items() if v[1] > 0 and v[2] > 0}}
        os.dup2(saved, 1)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
