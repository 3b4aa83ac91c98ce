"""sg_rb2d_state_serialize / deserialize from the device-resident rigidbody2d state.  The byte layout itself is checked on the CPU against
the reference's own RigidBody2DState::serialize (tests/test_rb2d_snapshot_cpu.py); what the GPU adds is where the arrays come from: the whole snapshot equals
what the reference's own RigidBody2DState writes for the same state -- as uploaded and after a resident step, with planes, kinematic circles, boxes, portals and
a Lees-Edwards offset -- and a context restored from it continues exactly like the original.
(Written with seconds of GPU time left in round 2: run on the B200 through profiles/lean_state_io_check.py -- profiles/lean_state_io_r2.log -- rather than through pytest.)"""
import os

import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sim(s, ctx, portals=None):
    import scisim_b200 as sb
    pp = sb.PlanarPortal.from_arrays(portals) if portals is not None else None
    st = sb.RigidBody2DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_of_body"], s["fixed"], s["M"], s["g"], s["plane_x"], s["plane_n"], planar_portals=pp)
    return sb.RigidBody2DSim(st, ctx=ctx)


def _context():
    import scisim_b200 as sb
    return sb.Context(0)


def _restore(blob, ctx):
    import scisim_b200 as sb
    return sb.RigidBody2DSim.deserializeState(blob, ctx)


@pytest.mark.parametrize("scene", ["circles_boxes", "kinematic_circles", "lees_edwards"])
def test_rb2d_snapshot_is_the_references_own_and_resumes(oracle, gpu_ctx, scene):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rb2d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    import scisim_b200 as sb
    from tests.reference_sim_binding import RefRB2DSim
    portals = None
    if scene == "circles_boxes":
        s = scenes.rb2d_random(3000, 171, nplanes=3)
    elif scene == "kinematic_circles":
        s = scenes.rb2d_random(3000, 172, kinds=("circle",), nfixed_frac=0.15, nplanes=2)
    else:
        s = scenes.rb2d_periodic(3000, 173, axes="xy", lees_edwards=0.7)
        portals = s["portals"]
    umap = sb.SymplecticEulerMap() if scene != "kinematic_circles" else sb.VerletMap()
    sim = _sim(s, gpu_ctx, portals)
    ref = RefRB2DSim(s, portals)
    sim.upload(s["q"], s["v"])
    assert sim.serializeState(which=0) == ref.serialize_state()
    with pytest.raises(sb.SciSimB200Error):
        sim.serializeState(which=1)          # nothing has flowed yet
    t = 3 * s["dt"]
    if portals is not None:
        sim.updatePeriodicBoundaryConditionsStartOfStep(3, s["dt"])
        ref.update_portals(t)
    c0 = sim.step(umap, s["dt"])
    q1, v1, _ = sim.fetch()
    blob = sim.serializeState(which=1)
    ref.set_state(q1, v1)
    theirs = ref.serialize_state()
    assert len(blob) == len(theirs) and blob == theirs
    # the reference reads the product's bytes back
    assert RefRB2DSim.from_snapshot(blob).serialize_state() == blob
    # resume in a fresh context: same snapshot back, same next step (the Lees-Edwards offset travels in the snapshot)
    ctx2 = _context()
    sim2 = _restore(blob, ctx2)
    assert sim2.nqdofs() == sim.nqdofs()
    assert sim2.serializeState(which=0) == blob
    sim.upload(q1, v1)
    c1 = sim.step(umap, s["dt"])
    c2 = sim2.step(umap, s["dt"])
    assert c1 == c2 and c1[1] > 0 and c0[1] > 0
    qa, va, aa = sim.fetch()
    qb, vb, ab = sim2.fetch()
    assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    for k in ("type", "i", "j", "aux", "n", "p"):
        assert np.array_equal(getattr(aa, k), getattr(ab, k)), k
    assert np.array_equal(aa.depth, ab.depth, equal_nan=True)
    ctx2.close()


def test_rb2d_snapshot_refusals(gpu_ctx):
    import scisim_b200 as sb
    s = scenes.rb2d_random(200, 174, nplanes=1)
    sim = _sim(s, gpu_ctx)
    sim.upload(s["q"], s["v"])
    blob = sim.serializeState(which=0)
    ctx2 = _context()
    with pytest.raises(sb.SciSimB200Error):
        _restore(blob[: len(blob) // 2], ctx2)
    with pytest.raises(sb.SciSimB200Error):
        _restore(blob[:-1], ctx2)
    ctx2.close()
