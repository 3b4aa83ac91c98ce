"""HISTORICAL (the knobs SG_BP_EMIT_AHEAD / SG_BP_CONTACT_PF were removed after this run; result: profiles/prefetch_ab_r2.jsonl).  Experiment (round 2, last session): L2 prefetch of the next gather hop in pass 2 (sg_bp_emit: mask block + own-row order words of the body `ahead` indices on) and
pass 3 (sg_bp_contacts: the next work item's partner record).  One process, the scene generated once per size; the knobs are read at every launch.
  python profiles/prefetch_ab.py [--sizes 2097152,16777216]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scisim_b200 as sb
from scisim_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="2097152,16777216")
args = ap.parse_args()
VARIANTS = [("base", {}), ("contact_pf", {"SG_BP_CONTACT_PF": "1"}), ("emit_64k", {"SG_BP_EMIT_AHEAD": "65536"}), ("emit_256k", {"SG_BP_EMIT_AHEAD": "262144"}),
            ("emit_1m", {"SG_BP_EMIT_AHEAD": "1048576"}), ("both_256k", {"SG_BP_CONTACT_PF": "1", "SG_BP_EMIT_AHEAD": "262144"})]
for n in [int(x) for x in args.sizes.split(",")]:
    s = scenes.ball2d_gas(n=n)
    steps = 10 if n <= (1 << 22) else 4
    ctx = sb.Context(0)
    sim = sb.Ball2DSim(sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"]), ctx=ctx)
    sim.upload(s["q"], s["v"])
    ref = None
    for name, env in VARIANTS:
        for k in ("SG_BP_CONTACT_PF", "SG_BP_EMIT_AHEAD"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for _ in range(3):
            ctx.flush_l2(); r = sim.step(sb.VerletMap(), s["dt"])
        ms = []
        for _ in range(steps):
            ctx.flush_l2(); ctx.timer_begin(); r = sim.step(sb.VerletMap(), s["dt"]); ms.append(ctx.timer_end())
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(steps):
            ctx.flush_l2(); sim.step(sb.VerletMap(), s["dt"])
        prof = ctx.profile(); ctx.profile_enable(False)
        ref = ref or r
        print(json.dumps({"bodies": n, "variant": name, "ms_per_step": round(float(np.mean(ms)), 4), "bp_emit_us": round(1e3 * prof["bp_emit"][1] / steps, 1),
                          "bp_contacts_us": round(1e3 * prof["bp_contacts"][1] / steps, 1), "same_counts": tuple(r) == tuple(ref)}), flush=True)
    ctx.close()
