"""rigidbody3d state I/O with triangle meshes (RigidBodyTriangleMesh.cpp:215-232): a mesh's record -- its whole input file -- is attached by the
caller (sg_rb3d_set_mesh_snapshot; here the reference's own RigidBodyTriangleMesh::serialize, mesh by mesh), and sg_rb3d_state_serialize writes, from the
device-resident state, the bytes the reference's own RigidBody3DState::serialize writes for the same scene; a context restored from them (the meshes re-added from
their records) steps exactly like the original.  The byte layout and the record parser are checked on the CPU (tests/test_rb3d_snapshot_cpu.py::test_mesh_snapshots).
(Written with seconds of GPU time left in round 2: run on the B200 through profiles/lean_state_io_check.py -- profiles/lean_state_io_r2.log -- rather than through pytest.)"""
import os

import numpy as np
import pytest

from scisim_b200 import scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sim(s, ctx):
    from tests.test_rb3d_gpu import make_sim
    return make_sim(s, ctx)


def _context():
    import scisim_b200 as sb
    return sb.Context(0)


def _restore(blob, ctx):
    import scisim_b200 as sb
    return sb.RigidBody3DSim.deserializeState(blob, ctx)


@pytest.mark.parametrize("scene", ["meshes", "mixed"])
def test_rb3d_mesh_snapshot_is_the_references_own_and_resumes(oracle, gpu_ctx, scene):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rb3d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    import scisim_b200 as sb
    from tests.reference_sim_binding import RefRB3DSim
    s = scenes.rb3d_random_meshes(60, 191, nplanes=2) if scene == "meshes" else scenes.rb3d_mixed_segregated(60, 192)
    n = s["geo_of_body"].shape[0]
    sim = _sim(s, gpu_ctx)
    ref = RefRB3DSim(s)
    sim.upload(s["q"], s["v"])
    with pytest.raises(sb.SciSimB200Error):
        sim.serializeState(which=0, m_updated=False)     # the meshes' records have not been attached
    for k in range(len(s["meshes"])):
        sim.setMeshSnapshot(k, ref.mesh_record(k))
    with pytest.raises(sb.SciSimB200Error):
        sim.setMeshSnapshot(0, ref.mesh_record(0)[:-8])   # not a whole record
    mine = sim.serializeState(which=0, m_updated=False)
    theirs = ref.serialize_state()
    assert len(mine) == len(theirs) and mine == theirs
    sim.step(sb.DMVMap(), s["dt"])
    q1, v1, a = sim.fetch()
    blob = sim.serializeState(which=1)
    assert blob == ref.serialize_state(q1, v1, update=True)
    # resume in a fresh context: the meshes come from their records; same snapshot back, same next step
    ctx2 = _context()
    sim2 = _restore(blob, ctx2)
    assert sim2.serializeState(which=0, m_updated=True) == blob
    sim.updateMandMinv()
    sim.upload(q1, v1)
    c1 = sim.step(sb.DMVMap(), s["dt"])
    c2 = sim2.step(sb.DMVMap(), s["dt"])
    assert c1 == c2 and c1[1] > 0
    qa, va, aa = sim.fetch()
    qb, vb, ab = sim2.fetch()
    assert np.array_equal(qa, qb) and np.array_equal(va, vb)
    for k in ("type", "i", "j", "aux", "n", "p"):
        assert np.array_equal(getattr(aa, k), getattr(ab, k)), k
    assert np.array_equal(aa.depth, ab.depth, equal_nan=True)
    assert int((aa.type == 12).sum()) > 0 or scene == "mixed"   # SG_BODY_BODY: mesh-mesh contacts
    ctx2.close()
