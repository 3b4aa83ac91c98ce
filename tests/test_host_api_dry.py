"""CPU-only: the Python mirror (scisim_b200/host_api.py) driven against a recording stand-in for the library.

No arithmetic happens here and nothing is claimed about the kernels: each call the sims make is converted with the REAL ctypes
argtypes of libscisim_b200.so (scisim_b200/_lib.py, themselves checked against the header in tests/test_abi.py), so a wrong
argument count, a Python object ctypes cannot convert, or a mis-shaped result view fails here rather than on the GPU box.  The
portal set-ups go through the very make_sim helpers of the GPU tests.
"""
import ctypes as C

import numpy as np
import pytest

from scisim_b200 import scenes


class RecordingLib:
    """Same attribute surface as the bound library; every call is type-converted, recorded and answered with 0 (SG_OK)."""

    def __init__(self):
        from scisim_b200 import _lib
        self.real = _lib.load()
        self.calls = []
        self.keep = []
        self.n_tele = 3
        self.dim = 2

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        argtypes = fn.argtypes

        def call(*args):
            assert len(args) == len(argtypes), "%s: %d arguments for %d parameters" % (name, len(args), len(argtypes))
            for a, t in zip(args, argtypes):
                t.from_param(a)  # raises ctypes.ArgumentError / TypeError as the real call would
            self.calls.append(name)
            if name.endswith("_teleported"):
                self.fill_teleported(args[1]._obj)
            if name == "sg_rb3d_add_mesh":
                args[-1]._obj.value = len([c for c in self.calls if c == name]) - 1
            return 0

        return call

    def fill_teleported(self, t):
        from scisim_b200 import _lib
        assert isinstance(t, _lib.SgTeleported)
        nt, nb = self.n_tele, 4
        u = lambda n: np.arange(n, dtype=np.uint32)
        d = lambda n: np.arange(n * self.dim, dtype=np.float64)
        arrs = dict(box_body=u(nb), box_portal=u(nb) + 10, portal0=u(nt), portal1=u(nt) + 1, x0=d(nt), x1=d(nt) + 0.5)
        if self.dim == 2:
            arrs.update(kick=np.zeros(2 * nt), delta0=np.ones(2 * nt), delta1=-np.ones(2 * nt))
        self.keep.append(arrs)
        t.n_boxes, t.n_regular, t.n_teleported = nb, 7, nt
        for k, a in arrs.items():
            setattr(t, k, a.ctypes.data_as(type(getattr(t, k))))


class DryContext:
    def __init__(self):
        self.lib = RecordingLib()
        self.h = C.c_void_p(1)
        self.device = 0

    def check(self, rc):
        assert rc == 0


@pytest.fixture
def dry():
    return DryContext()


def test_ball2d_portal_calls_convert(dry):
    import scisim_b200 as sb
    s = scenes.ball2d_periodic(50, 1, lees_edwards=0.5)
    st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"], planar_portals=sb.PlanarPortal.from_arrays(s["portals"]))
    sim = sb.Ball2DSim(st, ctx=dry)
    assert "sg_ball2d_set_portals" in dry.lib.calls
    dx = sim.updatePeriodicBoundaryConditionsStartOfStep(1, 0.01)
    assert dx.shape == (len(st.planar_portals),)
    q, v = sim.enforcePeriodicBoundaryConditions(s["q"], s["v"])
    assert q.shape == s["q"].shape and v.shape == s["v"].shape
    got = sim.computeActiveSet(s["q"], s["q"], s["v"])
    assert got.n_active == 0
    t = sim.teleported()
    assert t.n_regular == 7 and t.x0.shape == (3, 2) and t.kick.shape == (3, 2) and t.box_portal.tolist() == [10, 11, 12, 13]


def test_rb2d_portal_calls_convert(dry):
    import scisim_b200 as sb
    from tests.test_zz_rb2d_portals_gpu import make_sim
    s = scenes.rb2d_periodic(60, 2, side=5.0, lees_edwards=0.7, t=1.3, oblique=True)
    sim = make_sim(s, dry)
    assert "sg_rb2d_set_portals" in dry.lib.calls and sim.name() == "rigid_body_2d"
    dx = sim.updatePeriodicBoundaryConditionsStartOfStep(1, s["t"])
    assert dx.shape == (len(s["portals"]["v"]),)
    q1, v1 = sb.VerletMap().flow(s["q"], s["v"], sim, 1, s["dt"])
    assert q1.shape == s["q"].shape and v1.shape == s["v"].shape
    assert sim.computeActiveSet(s["q"], q1).n_active == 0
    assert sim.computeActiveSet(s["q"], q1, resident=True).n_active == 0
    t = sim.teleported()
    assert t.delta0.shape == (3, 2) and np.all(t.delta1 == -1.0) and t.portal1.tolist() == [1, 2, 3]
    q, v = sim.enforcePeriodicBoundaryConditions(s["q"], s["v"])
    assert np.array_equal(q, s["q"]) and np.array_equal(v, s["v"])  # the stand-in computes nothing
    # no portals: the sim still tells the library so
    st = sb.RigidBody2DState(s["geo_type"], s["geo_r"], s["geo_half"], s["geo_of_body"], s["fixed"], s["M"], s["g"], s["plane_x"], s["plane_n"])
    n0 = dry.lib.calls.count("sg_rb2d_set_portals")
    sb.RigidBody2DSim(st, ctx=dry)
    assert dry.lib.calls.count("sg_rb2d_set_portals") == n0 + 1


def test_rb3d_portal_calls_convert(dry):
    import scisim_b200 as sb
    from tests.test_zz_rb3d_portals_gpu import make_sim
    dry.lib.dim = 3
    s = scenes.rb3d_periodic_spheres(40, 3, side=4.0)
    sim = make_sim(s, dry)
    assert "sg_rb3d_set_portals" in dry.lib.calls and sim.name() == "rigid_body_3d"
    assert sim.computeActiveSet(s["q"], s["q"]).n_active == 0
    t = sim.teleported()
    assert t.x0.shape == (3, 3) and t.kick is None and t.delta0 is None and t.n_teleported == 3
    q = sim.enforcePeriodicBoundaryConditions(s["q"])
    assert q.shape == s["q"].shape
    mults = np.array([p.multiplier for p in sim.state.planar_portals])
    assert mults.dtype == np.int32 and mults.shape == (len(s["portals"]["mult"]), 3)


def test_rb3d_update_m_and_minv_calls_convert(dry):
    import scisim_b200 as sb
    from scisim_b200._lib import SG_MAP_M_UPDATED
    from tests.test_rb3d_gpu import make_sim
    s = scenes.rb3d_random_boxes(20, 1)
    sim = make_sim(s, dry)
    seen = []
    real = dry.lib.__getattr__("sg_rb3d_flow")
    dry.lib.__dict__["sg_rb3d_flow"] = lambda *a: (seen.append(a[1]), real(*a))[1]
    sb.DMVMap().flow(s["q"], s["v"], sim, 1, s["dt"])
    I, Ii = sim.updateMandMinv(s["q"])
    assert I.shape == (180,) and Ii.shape == (180,)
    sim.updateMandMinv()
    sb.DMVMap().flow(s["q"], s["v"], sim, 2, s["dt"])
    assert seen == [3, 3 | SG_MAP_M_UPDATED]  # SG_MAP_DMV, then with the flag once updateMandMinv has run


def test_rb2d_resident_calls_convert(dry):
    import scisim_b200 as sb
    from tests.test_rb2d_gpu import make_sim
    s = scenes.rb2d_random(30, 2)
    sim = make_sim(s, dry)
    sim.upload(s["q"], s["v"])
    assert sim.step(sb.VerletMap(), s["dt"]) == (0, 0)
    q1, v1, a = sim.fetch()
    assert q1.shape == s["q"].shape and v1.shape == s["v"].shape and a.n_active == 0
    q1, v1, a = sim.fetch(want_state=False)
    assert q1 is None and v1 is None
    assert [c for c in dry.lib.calls if c.startswith("sg_rb2d_") and c.split("_")[-1] in ("upload", "step", "fetch")] == ["sg_rb2d_upload", "sg_rb2d_step", "sg_rb2d_fetch", "sg_rb2d_fetch"]


def test_a_wrong_argument_is_caught(dry):
    with pytest.raises(AssertionError):
        dry.lib.sg_rb3d_enforce_portals(dry.h, None, None)
    with pytest.raises((C.ArgumentError, TypeError)):
        dry.lib.sg_rb2d_update_portals(dry.h, "soon", None)


def test_rb2d_state_io_calls_convert(dry):
    """serializeState / deserializeState of the rigidbody2d mirror: the size-query call fills *bytes, the restore sizes its host-side mirror from the blob."""
    import scisim_b200 as sb
    from tests.test_zzz_rb2d_state_io_gpu import _sim
    s = scenes.rb2d_periodic(40, 4, axes="xy", lees_edwards=0.3)
    sim = _sim(s, dry, s["portals"])
    real = dry.lib.__getattr__("sg_rb2d_state_serialize")
    dry.lib.__dict__["sg_rb2d_state_serialize"] = lambda *a: (setattr(a[4]._obj, "value", 24), real(*a))[1]
    blob = sim.serializeState(which=0)
    assert isinstance(blob, bytes) and len(blob) == 24
    assert dry.lib.calls.count("sg_rb2d_state_serialize") == 2
    # a snapshot laid out as sg_rb2d_snapshot.h writes it: 5 bodies, 2 geometries, 1 force, 3 planes, 2 portals
    n, nq = 5, 15
    i64, u64 = (lambda v: np.int64(v).tobytes()), (lambda v: np.uint64(v).tobytes())
    sparse = i64(nq) * 3 + bytes(4 * nq) + bytes(4 * (nq + 1)) + bytes(8 * nq)
    layout = (i64(nq) + bytes(8 * nq)) * 2 + sparse * 2 + u64(n) + bytes(n) + i64(n) + bytes(4 * n) + u64(2) + np.int32(0).tobytes() + bytes(8) + np.int32(1).tobytes() + bytes(16) \
        + u64(1) + bytes(4 + 16) + u64(3) + bytes(3 * 72) + u64(2) + bytes(2 * (2 * 72 + 24))
    assert sb.host_api.rb2d_snapshot_counts(layout) == (5, 2)
    sim2 = sb.RigidBody2DSim.deserializeState(layout, dry)
    assert sim2.nqdofs() == 15 and len(sim2.state.planar_portals) == 2 and "sg_rb2d_state_deserialize" in dry.lib.calls
    assert sim2.updatePeriodicBoundaryConditionsStartOfStep(2, 0.01).shape == (2,)
    sim2.upload(np.zeros(15), np.zeros(15))
    assert sim2.step(sb.SymplecticEulerMap(), 0.01) == (0, 0)


def test_rb3d_mesh_snapshot_calls_convert(dry):
    """RigidBody3DSim numbers its meshes through the indices sg_rb3d_add_mesh returns -- a context that already holds meshes hands out later ones -- and
    setMeshSnapshot addresses them the same way."""
    import scisim_b200 as sb
    from tests.test_rb3d_gpu import make_sim
    dry.lib.dim = 3
    s = scenes.rb3d_random_meshes(6, 5, nplanes=1)
    make_sim(s, dry)                 # an earlier sim on the same context: meshes 0 and 1
    seen = []
    real = dry.lib.__getattr__("sg_rb3d_set_geometry")
    dry.lib.__dict__["sg_rb3d_set_geometry"] = lambda *a: (seen.append(np.ctypeslib.as_array(C.cast(a[5], C.POINTER(C.c_uint32)), shape=(int(a[1]),)).copy()), real(*a))[1]
    sim = make_sim(s, dry)
    assert sim.mesh_index == [2, 3] and seen[-1].tolist() == [2, 3]
    which = []
    real2 = dry.lib.__getattr__("sg_rb3d_set_mesh_snapshot")
    dry.lib.__dict__["sg_rb3d_set_mesh_snapshot"] = lambda *a: (which.append((int(a[1]), int(a[3]))), real2(*a))[1]
    sim.setMeshSnapshot(1, b"\x03" + bytes(99))
    assert which == [(3, 100)]
