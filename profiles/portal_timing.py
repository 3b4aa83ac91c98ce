"""Resident-step timing of the ball2d portal path on ONE B200 (run under gpurun):  python profiles/portal_timing.py [n]
Periodic box, plain portal pair in x + Lees-Edwards pair in y, Verlet.  Prints one JSON line: pairs/s, ms/step, launches per
step and per-kernel microseconds (CUDA events inside the library)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scisim_b200 as sb
from scisim_b200 import scenes


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    steps, warmup = 5, 2
    s = scenes.ball2d_periodic(n, 7, axes="xy", lees_edwards=0.4, t=0.9)
    ctx = sb.Context(0)
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"], planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
    sim = sb.Ball2DSim(st, ctx=ctx)
    sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    sim.upload(s["q"], s["v"])
    umap = sb.VerletMap()
    for _ in range(warmup):
        ctx.flush_l2(); r = sim.step(umap, s["dt"])
    ms = []
    l0 = ctx.launch_count()
    for _ in range(steps):
        ctx.flush_l2(); ctx.timer_begin(); r = sim.step(umap, s["dt"]); ms.append(ctx.timer_end())
    launches = (ctx.launch_count() - l0) / steps
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(steps):
        ctx.flush_l2(); sim.step(umap, s["dt"])
    prof = ctx.profile(); ctx.profile_enable(False)
    tele = sim.teleported()
    pc, pa = r
    print(json.dumps({"workload": "ball2d periodic box, %d balls, x: planar portal, y: Lees-Edwards portal, Verlet" % n, "candidates": int(pc), "active": int(pa),
                      "teleported_boxes": int(tele.box_body.shape[0]), "teleported_contacts": int(tele.n_teleported), "ms_per_step": round(float(np.mean(ms)), 4),
                      "pairs_per_s": (pc + pa) / (float(np.mean(ms)) * 1e-3), "launches_per_step": launches,
                      "kernels_us": {k: round(1e3 * v[1] / steps, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}))


if __name__ == "__main__":
    main()
