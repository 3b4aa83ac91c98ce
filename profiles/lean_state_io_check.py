"""The six state-I/O GPU tests added at the end of round 2, run without pytest / torch so that they fit the few seconds of GPU time the round had left:
   python profiles/lean_state_io_check.py   (writes gpurun_out/lean_state_io.log line by line)"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.makedirs("gpurun_out", exist_ok=True)
log = open("gpurun_out/lean_state_io.log", "w")


def say(s):
    log.write(s + "\n"); log.flush(); os.fsync(log.fileno())
    print(s, flush=True)


t0 = time.time()
import scisim_b200 as sb
from tests import oracle_binding
import tests.test_zzz_rb2d_state_io_gpu as a
import tests.test_zzz_rb3d_mesh_state_io_gpu as b
oracle = oracle_binding.load()
ctx = sb.Context(0)
say("context %.2f s" % (time.time() - t0))
jobs = [("rb2d " + s, lambda s=s: a.test_rb2d_snapshot_is_the_references_own_and_resumes(oracle, ctx, s)) for s in ("circles_boxes", "kinematic_circles", "lees_edwards")]
jobs.append(("rb2d refusals", lambda: a.test_rb2d_snapshot_refusals(ctx)))
jobs += [("rb3d " + s, lambda s=s: b.test_rb3d_mesh_snapshot_is_the_references_own_and_resumes(oracle, ctx, s)) for s in ("meshes", "mixed")]
for name, job in jobs:
    try:
        job()
        say("PASS %s %.2f s" % (name, time.time() - t0))
    except BaseException as e:  # noqa: BLE001 -- pytest.skip included
        say("FAIL %s %.2f s: %r\n%s" % (name, time.time() - t0, e, traceback.format_exc()[-1500:]))
say("done")
