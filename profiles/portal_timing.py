"""Timing of the portal paths on ONE B200 (run under gpurun):  python profiles/portal_timing.py [ball2d|rb2d|rb3d] [n]

ball2d: periodic box, plain portal pair in x + Lees-Edwards pair in y, Verlet, resident step (sg_ball2d_step).
rb3d:   periodic box of spheres, portals in x and z, DMV, resident step (sg_rb3d_step).
rb2d:   periodic box of circles, x plain + y Lees-Edwards, Verlet, resident step (sg_rb2d_step).
Prints one JSON line: pairs/s, ms/step, launches per step and per-kernel microseconds (CUDA events inside the library)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scisim_b200 as sb
from scisim_b200 import scenes


def build(system, n, ctx):
    if system == "ball2d":
        s = scenes.ball2d_periodic(n, 7, axes="xy", lees_edwards=0.4, t=0.9)
        st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"], planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
        sim = sb.Ball2DSim(st, ctx=ctx)
        sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
        return s, sim, sb.VerletMap(), "ball2d periodic box, %d balls, x: planar portal, y: Lees-Edwards portal, Verlet, resident step" % n
    if system == "rb3d":
        s = scenes.rb3d_periodic_spheres(n, 7, axes="xz")
        st = sb.RigidBody3DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_mesh"], [], s["geo_of_body"], s["fixed"], s["m"], s["I0"], s["g"], s["plane_x"], s["plane_n"],
                                 planar_portals=sb.PlanarPortal3D.from_arrays(s["portals"]))
        return s, sb.RigidBody3DSim(st, ctx=ctx), sb.DMVMap(), "rigidbody3d periodic box, %d spheres, portals in x and z, DMV, resident step" % n
    s = scenes.rb2d_periodic(n, 7, lees_edwards=0.4, t=0.9)
    st = sb.RigidBody2DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_of_body"], s["fixed"], s["M"], s["g"], s["plane_x"], s["plane_n"],
                             planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
    sim = sb.RigidBody2DSim(st, ctx=ctx)
    sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    return s, sim, sb.VerletMap(), "rigidbody2d periodic box, %d circles, x: planar portal, y: Lees-Edwards portal, Verlet, resident step" % n


def main():
    system = sys.argv[1] if len(sys.argv) > 1 and sys.argv[1] in ("ball2d", "rb2d", "rb3d") else "ball2d"
    nums = [a for a in sys.argv[1:] if a.isdigit()]
    n = int(nums[0]) if nums else 1 << 20
    steps, warmup = 5, 2
    ctx = sb.Context(0)
    s, sim, umap, workload = build(system, n, ctx)
    sim.upload(s["q"], s["v"])

    def step():
        return sim.step(umap, s["dt"])
    for _ in range(warmup):
        ctx.flush_l2(); r = step()
    ms = []
    l0 = ctx.launch_count()
    for _ in range(steps):
        ctx.flush_l2(); ctx.timer_begin(); r = step(); ms.append(ctx.timer_end())
    launches = (ctx.launch_count() - l0) / steps
    ctx.profile_enable(True); ctx.profile_reset()
    for _ in range(steps):
        ctx.flush_l2(); step()
    prof = ctx.profile(); ctx.profile_enable(False)
    tele = sim.teleported()
    pc, pa = r
    print(json.dumps({"workload": workload, "candidates": int(pc), "active": int(pa),
                      "teleported_boxes": int(tele.box_body.shape[0]), "teleported_contacts": int(tele.n_teleported), "ms_per_step": round(float(np.mean(ms)), 4),
                      "timed": "CUDA events around the resident step",
                      "pairs_per_s": (pc + pa) / (float(np.mean(ms)) * 1e-3), "launches_per_step": launches,
                      "kernels_us": {k: round(1e3 * v[1] / steps, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}))


if __name__ == "__main__":
    main()
