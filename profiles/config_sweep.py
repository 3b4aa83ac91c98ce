"""Resident-step timing of the other BASELINE configs on ONE B200 (run under gpurun):  python profiles/config_sweep.py
Prints one JSON line per workload: pairs/s, ms/step, per-kernel us (CUDA events inside the library)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scisim_b200 as sb
from scisim_b200 import scenes


def timed(ctx, step, steps=10, warmup=3):
    for _ in range(warmup):
        ctx.flush_l2(); r = step()
    ms = []
    for _ in range(steps):
        ctx.flush_l2(); ctx.timer_begin(); r = step(); ms.append(ctx.timer_end())
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(steps):
        ctx.flush_l2(); step()
    prof = ctx.profile(); ctx.profile_enable(False)
    return r, float(np.mean(ms)), {k: round(1e3 * v[1] / steps, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}


def main():
    ctx = sb.Context(0)
    out = []
    # config 3 slice: polydisperse gas, 2M balls per GPU (16M over 8 GPUs), Verlet
    s = scenes.ball2d_gas(n=1 << 21)
    sim = sb.Ball2DSim(sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"]), ctx=ctx)
    sim.upload(s["q"], s["v"])
    (pc, pa), ms, k = timed(ctx, lambda: sim.step(sb.VerletMap(), s["dt"]))
    out.append({"workload": "config 3 slice: 2M-ball polydisperse gas (r 0.25..1), Verlet", "candidates": int(pc), "active": int(pa), "ms_per_step": round(ms, 4), "pairs_per_s": (pc + pa) / (ms * 1e-3), "kernels_us": k})
    del sim
    # config 4: 4M-sphere box drop, rb3d, DMV map (fused sphere route)
    s = scenes.rb3d_sphere_lattice(160)
    from tests.test_rb3d_gpu import make_sim
    sim = make_sim(s, ctx)
    sim.upload(s["q"], s["v"])
    (pc, pa), ms, k = timed(ctx, lambda: sim.step(sb.DMVMap(), s["dt"]), steps=5)
    out.append({"workload": "config 4: %d-sphere lattice drop, rb3d, DMV" % s["geo_of_body"].shape[0], "candidates": int(pc), "active": int(pa), "ms_per_step": round(ms, 4), "pairs_per_s": (pc + pa) / (ms * 1e-3), "kernels_us": k})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
