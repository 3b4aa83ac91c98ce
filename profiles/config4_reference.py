"""CPU figure for BASELINE configs[3] (rigidbody3d sphere box drop, split_ham) from the reference's OWN code: RigidBody3DSim::flow( SplitHamMap ) +
RigidBody3DSim::computeActiveSet of rigidbody3d/RigidBody3DSim.cpp, compiled unchanged into oracle/_ref (oracle/Makefile.ref, oracle/ref_shims/ref_rb3d_sim.cpp),
one thread (the reference's path is single-threaded), on a bounded sample of the same generator (side^3 spheres; the full scene is 160^3).
  python profiles/config4_reference.py [--side 64] [--steps 3] [--warmup 1]
Pairs = candidate pairs (the oracle's count for the same state, untimed) + active contacts (the reference's own list), as in the GPU arm."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scisim_b200 import scenes
from tests import oracle_binding as ob
from tests.reference_sim_binding import RefRB3DSim, f64, vp

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()
s = scenes.rb3d_sphere_lattice(args.side, args.side, args.side)
n = args.side ** 3
ref = RefRB3DSim(s)
f = ref.lib.ref_rb3d_sim_step_timed
f.restype = C.c_double
f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]
q0, v0 = f64(s["q"]), f64(s["v"])
assert s["dt"] == 1.0e-3
o = ob.RB3DOracle(s)
q1, _ = o.flow(2, q0, v0, s["dt"])
cand = o.active_set(q0, q1, "grid")["candidates"].shape[0]
times, na = [], C.c_uint64(0)
for it in range(args.warmup + args.steps):
    tf = C.c_double(0.0)
    t = float(f(ref.h, vp(q0), vp(v0), 2, 1, 1, 1000, C.byref(na), C.byref(tf)))
    if it >= args.warmup:
        times.append(t)
pairs = cand + int(na.value)
print(json.dumps({"impl": "reference", "workload": "configs[3] sample: %d^3 = %d spheres (same generator as the 160^3 scene), rigidbody3d, split_ham" % (args.side, n), "bodies": n,
                  "candidates": cand, "active": int(na.value), "ms_per_step": 1e3 * sum(times) / len(times), "pairs_per_s": pairs * len(times) / sum(times), "cores": 1,
                  "kind": "reference", "what": "the reference's own RigidBody3DSim::flow( SplitHamMap ) + RigidBody3DSim::computeActiveSet, compiled unchanged against oracle/eigen_standin, -O3 -DNDEBUG"}))
