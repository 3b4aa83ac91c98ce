"""HISTORICAL (the knob SG_BP_MINB was removed after this run; result: profiles/minb_ab_r2.jsonl).  Experiment (round 2, last session): resident CTAs per SM of the un-staged 2-D pass 1 (sg_bp_count_l1<P, 2, MINB>): 4 (64 registers), 5 (48, a few spills),
6 (40, more spills).  One process, the scene generated once per size, one context per variant (the knob SG_BP_MINB is read when a context first launches pass 1).
  python profiles/minb_ab.py [--sizes 2097152,16777216]
Prints one JSON line per (size, variant): ms/step and the us of bp_count; the candidate / contact counts must agree between variants."""
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
ap.add_argument("--variants", default="4,5,6")
args = ap.parse_args()
for n in [int(x) for x in args.sizes.split(",")]:
    s = scenes.ball2d_gas(n=n)
    steps = 10 if n <= (1 << 22) else 4
    ref = None
    for minb in [int(x) for x in args.variants.split(",")]:
        os.environ["SG_BP_MINB"] = str(minb)
        ctx = sb.Context(0)
        sim = sb.Ball2DSim(sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"]), ctx=ctx)
        sim.upload(s["q"], s["v"])
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
        print(json.dumps({"bodies": n, "minb": minb, "ms_per_step": round(float(np.mean(ms)), 4), "bp_count_us": round(1e3 * prof["bp_count"][1] / steps, 1),
                          "counts": [int(r[0]), int(r[1])], "same_counts": tuple(r) == tuple(ref)}), flush=True)
        del sim
        ctx.close()
