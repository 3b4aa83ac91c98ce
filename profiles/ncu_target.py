"""A few resident steps of one BASELINE config on ONE B200, as the target of an ncu capture or a plain timing run:
  python profiles/ncu_target.py --config 3 [--n 2097152] [--steps 2] [--warmup 2] [--time]
--config 2: 1M-ball pile; 3: polydisperse gas (n balls, default 2M = the per-GPU share of 16M on 8 GPUs); 4: rb3d sphere lattice (side
--side, default 160 => 4 096 000 spheres, split_ham); 5: mixed sphere/box/mesh scene.  --time prints one JSON line with ms/step and per-kernel us."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scisim_b200 as sb
from scisim_b200 import scenes


def build(cfg, ctx, args):
    if cfg == 2:
        s = scenes.ball2d_lattice(1000, 1000, seed=42, with_planes=True)
        sim = sb.Ball2DSim(sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"]), ctx=ctx)
        sim.upload(s["q"], s["v"])
        return (lambda: sim.step(sb.SymplecticEulerMap(), s["dt"])), "config 2: 1M-ball pile", s["r"].shape[0]
    if cfg == 3:
        s = scenes.ball2d_gas(n=args.n)
        sim = sb.Ball2DSim(sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"]), ctx=ctx)
        sim.upload(s["q"], s["v"])
        return (lambda: sim.step(sb.VerletMap(), s["dt"])), "config 3: %d-ball polydisperse gas, Verlet" % args.n, args.n
    if cfg == 4:
        s = scenes.rb3d_sphere_lattice(args.side, args.side, args.side)
        from tests.test_rb3d_gpu import make_sim
        sim = make_sim(s, ctx)
        sim.upload(s["q"], s["v"])
        umap = sb.SplitHamMap() if args.map == "split_ham" else sb.DMVMap()
        return (lambda: sim.step(umap, s["dt"])), "config 4: %d-sphere lattice drop, rb3d, %s" % (args.side ** 3, args.map), args.side ** 3
    if cfg == 5:
        s = scenes.rb3d_mixed_segregated(nper=args.nper)
        from tests.test_rb3d_gpu import make_sim
        sim = make_sim(s, ctx)
        sim.upload(s["q"], s["v"])
        return (lambda: sim.step(sb.DMVMap(), s["dt"])), "config 5: %d mixed sphere/box/mesh bodies, dmv" % s["geo_of_body"].shape[0], s["geo_of_body"].shape[0]
    if cfg == 6:
        # rigidbody2d: circles and rotated boxes (circle-circle CCD, circle-box, box-box, planes), symplectic Euler
        s = scenes.rb2d_random(args.n, 61, nfixed_frac=0.0, nplanes=3)
        from tests.test_rb2d_gpu import make_sim
        sim = make_sim(s, ctx)
        sim.upload(s["q"], s["v"])
        return (lambda: sim.step(sb.SymplecticEulerMap(), s["dt"])), "rigidbody2d: %d circles and boxes, symplectic Euler" % args.n, args.n
    if cfg == 7:
        # ball2d with portals: planar portal in x, Lees-Edwards in y (the portal branch of Ball2DSim::computeActiveSet)
        s = scenes.ball2d_periodic(args.n, 62, lees_edwards=0.7, t=0.3)
        from tests.test_portals_gpu import make_sim
        sim = make_sim(s, ctx)
        sim.updatePeriodicBoundaryConditionsStartOfStep(3, 0.1)
        sim.upload(s["q"], s["v"])
        return (lambda: sim.step(sb.SymplecticEulerMap(), s["dt"])), "ball2d with portals: %d balls, planar portal in x, Lees-Edwards in y" % args.n, args.n
    raise SystemExit("unknown config")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--n", type=int, default=1 << 21)
    ap.add_argument("--side", type=int, default=160)
    ap.add_argument("--nper", type=int, default=3334)
    ap.add_argument("--map", default="split_ham")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    ctx = sb.Context(0)
    step, name, n = build(args.config, ctx, args)
    for _ in range(args.warmup):
        ctx.flush_l2(); r = step()
    ms = []
    for _ in range(args.steps):
        ctx.flush_l2(); ctx.timer_begin(); r = step(); ms.append(ctx.timer_end())
    if args.time:
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(args.steps):
            ctx.flush_l2(); step()
        prof = ctx.profile(); ctx.profile_enable(False)
        pc, pa = int(r[0]), int(r[1])
        print(json.dumps({"workload": name, "bodies": n, "candidates": pc, "active": pa, "ms_per_step": round(float(np.mean(ms)), 4), "pairs_per_s": (pc + pa) / (float(np.mean(ms)) * 1e-3),
                          "kernels": {k: {"us": round(1e3 * v[1] / args.steps, 1), "alg_GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}))


if __name__ == "__main__":
    main()
