"""CPU: the product's RigidBody3DState snapshot writer / parser (scisim_b200/csrc/sg_rb3d_snapshot.h -- the header sg_rb3d.cu includes, compiled here for the
host through tests/rb3d_snapshot_harness.cpp) against the reference's OWN RigidBody3DState::serialize / deserialize (rigidbody3d/RigidBody3DState.cpp compiled
unchanged into oracle/_ref; the leaf serialisers of MathUtilities restated in the reference's byte layout): the same bytes for the same state, both mass-matrix
layouts (as constructed / after updateMandMinv), planes, cylinders, portals, kinematic bodies; parse -> write is the identity; the reference reads the
product's bytes back.  What the GPU adds (tests/test_state_io_gpu.py) is only where the arrays come from."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob
from tests.reference_sim_binding import RefRB3DSim, f64, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("snap") / "libsnap_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "rb3d_snapshot_harness.cpp")], check=True)
    lib = C.CDLL(out)
    V = C.c_void_p
    lib.snap_serialize.restype = C.c_uint64
    lib.snap_serialize.argtypes = [C.c_uint32] + [V] * 8 + [C.c_uint32] + [V] * 4 + [C.c_uint32, V, V, C.c_uint32, V, V, V, C.c_uint32, V, V, V, V, V, V, C.c_uint64, V, V]
    lib.snap_mesh_records.restype = C.c_int
    lib.snap_mesh_records.argtypes = [V, C.c_uint64, C.c_uint32, V, V, V, V]
    lib.snap_roundtrip.restype = C.c_int
    lib.snap_roundtrip.argtypes = [V, C.c_uint64, V, C.c_uint64, V, V, V]
    return lib


def _normalized(n):
    """StaticPlane's / StaticCylinder's constructors: n.normalized() with the 3-term squared norm ( a0 a0 + a1 a1 ) + a2 a2."""
    n = np.asarray(n, dtype=np.float64)
    z = (n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1]) + n[:, 2] * n[:, 2]
    return n / np.sqrt(z)[:, None]


def product_bytes(lib, s, q, v, blocks, portals=None, mesh_records=None):
    """mesh_records: per geometry, the bytes RigidBodyTriangleMesh::serialize writes for it (None / b"" for boxes and spheres)."""
    n = s["geo_of_body"].shape[0]
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
    I, Iinv = blocks
    gt = u32([int(t) if int(t) in (0, 3) else 1 for t in s["geo_type"]])
    px, pn = f64(s["plane_x"]).reshape(-1, 3), _normalized(f64(s["plane_n"]).reshape(-1, 3)) if len(s["plane_x"]) else np.zeros((0, 3))
    cyl = (f64(s["cyl_x"]), _normalized(s["cyl_axis"]), f64(s["cyl_r"])) if "cyl_r" in s and len(s["cyl_r"]) else (np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
    p = portals or {"plane_a_x": np.zeros((0, 3)), "plane_a_n": np.zeros((0, 3)), "plane_b_x": np.zeros((0, 3)), "plane_b_n": np.zeros((0, 3)), "mult": np.zeros((0, 3), np.int32)}
    pan = _normalized(p["plane_a_n"]) if len(p["plane_a_n"]) else np.zeros((0, 3))
    pbn = _normalized(p["plane_b_n"]) if len(p["plane_b_n"]) else np.zeros((0, 3))
    mult = np.ascontiguousarray(p["mult"], dtype=np.int32)
    k = [f64(q), f64(v), f64(s["m"]), f64(s["I0"]), f64(I), f64(Iinv), np.ascontiguousarray(s["fixed"], dtype=np.uint8), u32(s["geo_of_body"]), gt, f64(s["geo_r"]), f64(s["geo_half"]), f64(s["g"]),
         f64(px), f64(pn), f64(cyl[0]), f64(cyl[1]), f64(cyl[2]), f64(p["plane_a_x"]), f64(pan), f64(p["plane_b_x"]), f64(pbn)]
    args = [n] + [vp(a) for a in k[:8]] + [gt.shape[0]] + [vp(a) for a in k[8:12]] + [k[12].shape[0], vp(k[12]), vp(k[13]), k[16].shape[0], vp(k[14]), vp(k[15]), vp(k[16]),
                                                                                         mult.shape[0], vp(k[17]), vp(k[18]), vp(k[19]), vp(k[20]), vp(mult)]
    rec = (None, None)
    if mesh_records is not None:
        raws = [np.frombuffer(r or b"", dtype=np.uint8).copy() for r in mesh_records]
        ptrs = (C.c_void_p * len(raws))(*[r.ctypes.data if r.shape[0] else None for r in raws])
        sizes = np.array([r.shape[0] for r in raws], dtype=np.uint64)
        rec = (ptrs, vp(sizes))
    need = int(lib.snap_serialize(*args, None, 0, *rec))
    if need == 0:
        return None
    buf = np.zeros(need, dtype=np.uint8)
    assert int(lib.snap_serialize(*args, vp(buf), need, *rec)) == need
    return buf.tobytes()


def _transposed(blocks, n):
    return blocks.reshape(n, 3, 3).transpose(0, 2, 1).reshape(-1).copy()


@pytest.mark.parametrize("scene", ["spheres_cylinders", "boxes", "portals"])
def test_snapshot_bytes_equal_the_references(oracle, harness, scene):
    portals = None
    if scene == "spheres_cylinders":
        s = scenes.rb3d_random_spheres(500, 141, spin=True, nfixed_frac=0.2, nplanes=3)
        s["cyl_x"] = np.array([[0.1, -0.2, 0.3], [0.0, 0.0, 0.0]]); s["cyl_axis"] = np.array([[0.2, 3.0, -0.1], [1.0, 0.1, 0.05]]); s["cyl_r"] = np.array([30.0, 40.0])
    elif scene == "boxes":
        s = scenes.rb3d_random_boxes(400, 142, spin=True, nfixed_frac=0.0, nplanes=2)
    else:
        s = scenes.rb3d_periodic_spheres(300, 143, axes="xz")
        portals = s["portals"]
    n = s["geo_of_body"].shape[0]
    o = ob.RB3DOracle(s)
    ref = RefRB3DSim(s, portals)
    q0, v0 = f64(s["q"]), f64(s["v"])
    # as constructed: the world-space blocks transposed (formWorldSpaceMassMatrix)
    I, Ii = o.update_m_and_minv(q0)
    theirs = ref.serialize_state()
    mine = product_bytes(harness, s, q0, v0, (_transposed(I, n), _transposed(Ii, n)), portals)
    assert len(mine) == len(theirs) and mine == theirs
    # a running simulation: ( q1, v1 ) of a flow, blocks as updateMandMinv leaves them
    q1, v1 = o.flow(3, q0, v0, s["dt"])
    I, Ii = o.update_m_and_minv(q1)
    theirs = ref.serialize_state(q1, v1, update=True)
    mine = product_bytes(harness, s, q1, v1, (I, Ii), portals)
    assert mine == theirs
    # parse -> write is the identity, and nothing is left unread
    raw = np.frombuffer(theirs, dtype=np.uint8).copy()
    out = np.zeros(raw.shape[0], dtype=np.uint8)
    nb, nn, g = C.c_uint64(0), C.c_uint32(0), np.zeros(3)
    assert harness.snap_roundtrip(vp(raw), raw.shape[0], vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 0
    assert int(nb.value) == raw.shape[0] and int(nn.value) == n and np.array_equal(out, raw) and np.array_equal(g, f64(s["g"]))
    # the reference reads the product's bytes and writes them back unchanged; restored, it detects what the original detects
    again = RefRB3DSim.from_snapshot(mine, n)
    assert again.serialize_state() == mine
    q2, _ = o.flow(3, q1, v1, s["dt"])
    a, b = ref.active_set(q1, q2), again.active_set(q1, q2)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert a["type"].shape[0] > 10


def test_truncated_and_foreign_snapshots_are_refused(oracle, harness):
    s = scenes.rb3d_random_spheres(50, 144, nplanes=1)
    ref = RefRB3DSim(s)
    blob = np.frombuffer(ref.serialize_state(), dtype=np.uint8).copy()
    out = np.zeros(blob.shape[0], dtype=np.uint8)
    nb, nn, g = C.c_uint64(0), C.c_uint32(0), np.zeros(3)
    for cut in (3, 100, blob.shape[0] // 2, blob.shape[0] - 1):
        assert harness.snap_roundtrip(vp(blob), cut, vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1
    # a mesh without its record: the writer refuses
    m = scenes.rb3d_random_meshes(4, 145, nplanes=0)
    o = ob.RB3DOracle(m)
    I, Ii = o.update_m_and_minv(f64(m["q"]))
    assert product_bytes(harness, m, f64(m["q"]), f64(m["v"]), (I, Ii)) is None
    assert product_bytes(harness, m, f64(m["q"]), f64(m["v"]), (I, Ii), mesh_records=[b"", b""]) is None


@pytest.mark.parametrize("scene", ["meshes", "mixed"])
def test_mesh_snapshots(oracle, harness, scene):
    """Triangle meshes: the state snapshot holds each mesh's own record (RigidBodyTriangleMesh::serialize, RigidBodyTriangleMesh.cpp:215-232).  With the records a
    caller attaches (here: the reference's own, mesh by mesh) the product's bytes are the reference's; the parser finds every record's extent and the arrays
    sg_rb3d_add_mesh takes; the reference resumes from the product's bytes."""
    s = scenes.rb3d_random_meshes(30, 146, nplanes=2) if scene == "meshes" else scenes.rb3d_mixed_segregated(40, 147)
    n = s["geo_of_body"].shape[0]
    o = ob.RB3DOracle(s)
    ref = RefRB3DSim(s)
    records = [ref.mesh_record(int(s["geo_mesh"][k])) if int(s["geo_type"][k]) == 3 else b"" for k in range(len(s["geo_type"]))]
    assert sum(1 for r in records if r) >= 2 and all(r[0] == 3 for r in records if r)
    q0, v0 = f64(s["q"]), f64(s["v"])
    I, Ii = o.update_m_and_minv(q0)
    theirs = ref.serialize_state()
    mine = product_bytes(harness, s, q0, v0, (_transposed(I, n), _transposed(Ii, n)), mesh_records=records)
    assert mine is not None and len(mine) == len(theirs) and mine == theirs
    # what the parser extracts: record extents and array sizes, per geometry
    raw = np.frombuffer(theirs, dtype=np.uint8).copy()
    ngeo = len(records)
    got_n, nbytes, sizes, sdf_sum = C.c_uint32(0), np.zeros(ngeo, np.uint64), np.zeros(4 * ngeo, np.uint32), np.zeros(ngeo)
    assert harness.snap_mesh_records(vp(raw), raw.shape[0], ngeo, C.byref(got_n), vp(nbytes), vp(sizes), vp(sdf_sum)) == 0
    assert int(got_n.value) == ngeo
    for k in range(ngeo):
        assert int(nbytes[k]) == len(records[k])
        if records[k]:
            m = s["meshes"][int(s["geo_mesh"][k])]
            assert list(sizes[4 * k:4 * k + 4]) == [m["verts"].shape[0], m["samples"].shape[0], m["hull"].shape[0], int(np.prod(m["dims"]))]
            assert np.isclose(sdf_sum[k], f64(m["sdf"]).sum(), rtol=1e-9)
    # parse -> write is the identity (the records travel through), truncated mesh records are refused
    out = np.zeros(raw.shape[0], dtype=np.uint8)
    nb, nn, g = C.c_uint64(0), C.c_uint32(0), np.zeros(3)
    assert harness.snap_roundtrip(vp(raw), raw.shape[0], vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 0
    assert int(nb.value) == raw.shape[0] and np.array_equal(out, raw)
    rec0 = [r for r in records if r][0]
    first = theirs.index(rec0)
    for cut in (first + 5, first + 40, first + len(rec0) // 2, first + len(rec0) - 1):
        assert harness.snap_roundtrip(vp(raw), cut, vp(out), out.shape[0], C.byref(nb), C.byref(nn), vp(g)) == 1
    # the reference reads the product's bytes back and detects what the original detects
    again = RefRB3DSim.from_snapshot(mine, n)
    assert again.serialize_state() == mine
    q1, _ = o.flow(3, q0, v0, s["dt"])
    a, b = ref.active_set(q0, q1), again.active_set(q0, q1)
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k
