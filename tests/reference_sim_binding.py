"""TEST INFRASTRUCTURE: ctypes bindings of oracle/ref_shims/ref_*_sim.cpp -- the reference's OWN Ball2DSim / RigidBody2DSim / RigidBody3DSim, compiled
unchanged into oracle/_ref (oracle/Makefile.ref).  Used by tests/test_reference_sim_cpu.py and by bench.py's CPU arms (the reference arm times
Ball2DSim::flow + Ball2DSim::computeActiveSet through RefBall2DSim).  Never imported by the product."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def available(name):
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", name))


def _lib(name):
    if not available(name):
        import pytest
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    return C.CDLL(os.path.join(ROOT, "oracle", "_ref", name))


class RefBall2DSim:
    def __init__(self, s, portals=None):
        self.lib = lib = _lib("libref_ball2d.so")
        lib.ref_ball2d_sim_create.restype = C.c_void_p
        lib.ref_ball2d_sim_create.argtypes = [C.c_uint32] + [C.c_void_p] * 6 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p] * 6
        lib.ref_ball2d_sim_destroy.argtypes = [C.c_void_p]
        lib.ref_ball2d_sim_active_set.restype = C.c_uint64
        lib.ref_ball2d_sim_active_set.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + [C.c_void_p] * 6
        lib.ref_ball2d_sim_flow.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]
        lib.ref_ball2d_sim_set_state.argtypes = [C.c_void_p] * 3
        self.n = n = s["r"].shape[0]
        p = portals or {"plane_a_x": np.zeros((0, 2)), "plane_a_n": np.zeros((0, 2)), "plane_b_x": np.zeros((0, 2)), "plane_b_n": np.zeros((0, 2)), "v": np.zeros(0), "bounds": np.zeros(0)}
        k = [f64(s["q"]), f64(s["v"]), f64(s["m"]), f64(s["r"]), np.zeros(n, dtype=np.uint8), f64(s["g"]), f64(s["plane_x"]), f64(s["plane_n"]), f64(s["drum_x"]), f64(s["drum_r"]),
             f64(p["plane_a_x"]), f64(p["plane_a_n"]), f64(p["plane_b_x"]), f64(p["plane_b_n"]), f64(p["v"]), f64(p["bounds"])]
        self.h = lib.ref_ball2d_sim_create(n, vp(k[0]), vp(k[1]), vp(k[2]), vp(k[3]), vp(k[4]), vp(k[5]), k[6].shape[0], vp(k[6]), vp(k[7]), k[8].shape[0], vp(k[8]), vp(k[9]),
                                           k[14].shape[0], vp(k[10]), vp(k[11]), vp(k[12]), vp(k[13]), vp(k[14]), vp(k[15]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_ball2d_sim_destroy(self.h)
            self.h = None

    def active_set(self, q0, q1):
        q0, q1 = f64(q0), f64(q1)
        cap = 16 * self.n + 64
        out = {"type": np.zeros(cap, np.uint32), "i": np.zeros(cap, np.uint32), "j": np.zeros(cap, np.uint32), "n": np.zeros((cap, 2)), "p": np.zeros((cap, 2)), "depth": np.zeros(cap)}
        na = int(self.lib.ref_ball2d_sim_active_set(self.h, vp(q0), vp(q1), None, cap, vp(out["type"]), vp(out["i"]), vp(out["j"]), vp(out["n"]), vp(out["p"]), vp(out["depth"])))
        assert na <= cap
        return {k: v[:na] for k, v in out.items()}

    def flow(self, kind, iteration, dt_num, dt_den):
        q, v = np.zeros(2 * self.n), np.zeros(2 * self.n)
        self.lib.ref_ball2d_sim_flow(self.h, kind, iteration, dt_num, dt_den, vp(q), vp(v))
        return q, v

    def set_state(self, q, v):
        self.lib.ref_ball2d_sim_set_state(self.h, vp(f64(q)), vp(f64(v)))

    def serialize_state(self):
        """Ball2DState::serialize of the simulation's current state (the reference's binary snapshot)."""
        self.lib.ref_ball2d_sim_serialize_state.restype = C.c_uint64
        self.lib.ref_ball2d_sim_serialize_state.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        need = int(self.lib.ref_ball2d_sim_serialize_state(self.h, None, 0))
        buf = np.zeros(need, dtype=np.uint8)
        assert int(self.lib.ref_ball2d_sim_serialize_state(self.h, vp(buf), need)) == need
        return buf.tobytes()

    @classmethod
    def from_snapshot(cls, blob, n):
        """Ball2DState::deserialize of a snapshot into a fresh reference simulation."""
        self = cls.__new__(cls)
        self.lib = lib = _lib("libref_ball2d.so")
        lib.ref_ball2d_sim_from_snapshot.restype = C.c_void_p
        lib.ref_ball2d_sim_from_snapshot.argtypes = [C.c_void_p, C.c_uint64]
        lib.ref_ball2d_sim_destroy.argtypes = [C.c_void_p]
        lib.ref_ball2d_sim_active_set.restype = C.c_uint64
        lib.ref_ball2d_sim_active_set.argtypes = [C.c_void_p] * 4 + [C.c_uint64] + [C.c_void_p] * 6
        lib.ref_ball2d_sim_flow.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]
        lib.ref_ball2d_sim_set_state.argtypes = [C.c_void_p] * 3
        raw = np.frombuffer(blob, dtype=np.uint8).copy()
        self.n = n
        self.h = lib.ref_ball2d_sim_from_snapshot(vp(raw), raw.shape[0])
        return self


class RefRB2DSim:
    def __init__(self, s, portals=None):
        self.lib = lib = _lib("libref_rb2d.so")
        V = C.c_void_p
        lib.ref_rb2d_sim_create.restype = V
        lib.ref_rb2d_sim_create.argtypes = [C.c_uint32, V, V, V, V, V, C.c_uint32, V, V, V, V, C.c_uint32, V, V, C.c_uint32, V, V, V, V, V, V]
        lib.ref_rb2d_sim_destroy.argtypes = [V]
        lib.ref_rb2d_sim_active_set.restype = C.c_uint64
        lib.ref_rb2d_sim_active_set.argtypes = [V, V, V, C.c_uint64, V, V, V, V, V, V, V]
        lib.ref_rb2d_sim_flow.argtypes = [V, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, V, V]
        lib.ref_rb2d_sim_set_state.argtypes = [V, V, V]
        self.n = n = s["geo_of_body"].shape[0]
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        p = portals or {"plane_a_x": np.zeros((0, 2)), "plane_a_n": np.zeros((0, 2)), "plane_b_x": np.zeros((0, 2)), "plane_b_n": np.zeros((0, 2)), "v": np.zeros(0), "bounds": np.zeros(0)}
        k = [f64(s["q"]), f64(s["v"]), f64(s["M"]), np.ascontiguousarray(s["fixed"], dtype=np.uint8), u32(s["geo_of_body"]), u32(s["geo_type"]), f64(s["geo_r"]), f64(s["geo_half"]), f64(s["g"]),
             f64(s["plane_x"]), f64(s["plane_n"]), f64(p["plane_a_x"]), f64(p["plane_a_n"]), f64(p["plane_b_x"]), f64(p["plane_b_n"]), f64(p["v"]), f64(p["bounds"])]
        self.h = lib.ref_rb2d_sim_create(n, vp(k[0]), vp(k[1]), vp(k[2]), vp(k[3]), vp(k[4]), k[5].shape[0], vp(k[5]), vp(k[6]), vp(k[7]), vp(k[8]), k[9].shape[0], vp(k[9]), vp(k[10]),
                                         k[15].shape[0], vp(k[11]), vp(k[12]), vp(k[13]), vp(k[14]), vp(k[15]), vp(k[16]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_rb2d_sim_destroy(self.h)
            self.h = None

    def active_set(self, q0, q1, cap=None):
        q0, q1 = f64(q0), f64(q1)
        cap = cap or 32 * self.n + 1024
        out = {"type": np.zeros(cap, np.uint32), "i": np.zeros(cap, np.uint32), "j": np.zeros(cap, np.uint32), "static": np.zeros(cap, np.uint32),
               "n": np.zeros((cap, 2)), "p": np.zeros((cap, 2)), "depth": np.zeros(cap)}
        na = int(self.lib.ref_rb2d_sim_active_set(self.h, vp(q0), vp(q1), cap, vp(out["type"]), vp(out["i"]), vp(out["j"]), vp(out["static"]), vp(out["n"]), vp(out["p"]), vp(out["depth"])))
        if na > cap:
            return self.active_set(q0, q1, cap=na)
        return {k: v[:na] for k, v in out.items()}

    def flow(self, kind, iteration, dt_num, dt_den):
        q, v = np.zeros(3 * self.n), np.zeros(3 * self.n)
        self.lib.ref_rb2d_sim_flow(self.h, kind, iteration, dt_num, dt_den, vp(q), vp(v))
        return q, v

    def set_state(self, q, v):
        self.lib.ref_rb2d_sim_set_state(self.h, vp(f64(q)), vp(f64(v)))

    def get_state(self):
        q, v = np.zeros(3 * self.n), np.zeros(3 * self.n)
        self.lib.ref_rb2d_sim_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.ref_rb2d_sim_get_state(self.h, vp(q), vp(v))
        return q, v

    def update_portals(self, t):
        """PlanarPortal::updateMovingPortals( t ) on every portal, as RigidBody2DSim::flow does before the map."""
        self.lib.ref_rb2d_sim_update_portals.argtypes = [C.c_void_p, C.c_double]
        self.lib.ref_rb2d_sim_update_portals(self.h, float(t))

    def serialize_state(self):
        """RigidBody2DState::serialize of the simulation's current state (the reference's binary snapshot)."""
        f = self.lib.ref_rb2d_sim_serialize_state
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        need = int(f(self.h, None, 0))
        buf = np.zeros(need, dtype=np.uint8)
        assert int(f(self.h, vp(buf), need)) == need
        return buf.tobytes()

    @classmethod
    def from_snapshot(cls, blob):
        """RigidBody2DState::deserialize of a snapshot into a fresh reference simulation."""
        self = cls.__new__(cls)
        self.lib = lib = _lib("libref_rb2d.so")
        V = C.c_void_p
        lib.ref_rb2d_sim_from_snapshot.restype = V
        lib.ref_rb2d_sim_from_snapshot.argtypes = [V, C.c_uint64, V]
        lib.ref_rb2d_sim_destroy.argtypes = [V]
        lib.ref_rb2d_sim_active_set.restype = C.c_uint64
        lib.ref_rb2d_sim_active_set.argtypes = [V, V, V, C.c_uint64, V, V, V, V, V, V, V]
        lib.ref_rb2d_sim_flow.argtypes = [V, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, V, V]
        lib.ref_rb2d_sim_set_state.argtypes = [V, V, V]
        raw = np.frombuffer(blob, dtype=np.uint8).copy()
        n = C.c_uint32(0)
        self.h = lib.ref_rb2d_sim_from_snapshot(vp(raw), raw.shape[0], C.byref(n))
        self.n = int(n.value)
        return self


class RefRB3DSim:
    def __init__(self, s, portals=None):
        self.lib = lib = _lib("libref_rb3d.so")
        V = C.c_void_p
        lib.ref_rb3d_mesh_create.restype = V
        lib.ref_rb3d_mesh_create.argtypes = [C.c_uint32, V, C.c_uint32, V, C.c_uint32, V, V, V, V, V]
        lib.ref_rb3d_mesh_destroy.argtypes = [V]
        lib.ref_rb3d_sim_create.restype = V
        lib.ref_rb3d_sim_create.argtypes = [C.c_uint32, V, V, V, V, V, V, C.c_uint32, V, V, V, V, V, C.c_uint32, V, V, C.c_uint32, V, V, V, C.c_uint32, V, V, V, V, V]
        lib.ref_rb3d_sim_destroy.argtypes = [V]
        lib.ref_rb3d_sim_active_set.restype = C.c_uint64
        lib.ref_rb3d_sim_active_set.argtypes = [V, V, V, C.c_uint64, V, V, V, V, V, V, V]
        self.n = n = s["geo_of_body"].shape[0]
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        self.meshes = []
        for m in s["meshes"]:
            a = [f64(m[k]) for k in ("verts", "samples", "hull", "cell_delta", "origin", "sdf")]
            self.meshes.append(lib.ref_rb3d_mesh_create(a[0].shape[0], vp(a[0]), a[1].shape[0], vp(a[1]), a[2].shape[0], vp(a[2]), vp(a[3]), vp(u32(m["dims"])), vp(a[4]), vp(a[5])))
        ngeo = s["geo_type"].shape[0]
        handles = (C.c_void_p * max(1, ngeo))()
        for k in range(ngeo):
            handles[k] = self.meshes[int(s["geo_mesh"][k])] if int(s["geo_type"][k]) == 3 else None
        cyl = (f64(s["cyl_x"]), f64(s["cyl_axis"]), f64(s["cyl_r"])) if "cyl_r" in s and len(s["cyl_r"]) else (np.zeros((0, 3)), np.zeros((0, 3)), np.zeros(0))
        p = portals or {"plane_a_x": np.zeros((0, 3)), "plane_a_n": np.zeros((0, 3)), "plane_b_x": np.zeros((0, 3)), "plane_b_n": np.zeros((0, 3)), "mult": np.zeros((0, 3), np.int32)}
        mult = np.ascontiguousarray(p["mult"], dtype=np.int32)
        k = [f64(s["q"]), f64(s["v"]), f64(s["m"]), f64(s["I0"]), np.ascontiguousarray(s["fixed"], dtype=np.uint8), u32(s["geo_of_body"]), u32(s["geo_type"]), f64(s["geo_r"]),
             f64(s["geo_half"]), f64(s["g"]), f64(s["plane_x"]), f64(s["plane_n"]), f64(p["plane_a_x"]), f64(p["plane_a_n"]), f64(p["plane_b_x"]), f64(p["plane_b_n"])]
        self.h = lib.ref_rb3d_sim_create(n, vp(k[0]), vp(k[1]), vp(k[2]), vp(k[3]), vp(k[4]), vp(k[5]), ngeo, vp(k[6]), vp(k[7]), vp(k[8]), C.cast(handles, V), vp(k[9]),
                                         k[10].shape[0], vp(k[10]), vp(k[11]), cyl[2].shape[0], vp(cyl[0]), vp(cyl[1]), vp(cyl[2]),
                                         mult.shape[0], vp(k[12]), vp(k[13]), vp(k[14]), vp(k[15]), vp(mult))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ref_rb3d_sim_destroy(self.h)
            self.h = None
            for m in self.meshes:
                self.lib.ref_rb3d_mesh_destroy(m)

    def mesh_record(self, k):
        """RigidBodyTriangleMesh::serialize of mesh k: the mesh's own record of a state snapshot (what sg_rb3d_set_mesh_snapshot takes)."""
        f = self.lib.ref_rb3d_mesh_serialize
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        need = int(f(self.meshes[k], None, 0))
        buf = np.zeros(need, dtype=np.uint8)
        assert int(f(self.meshes[k], vp(buf), need)) == need
        return buf.tobytes()

    def serialize_state(self, q=None, v=None, update=False):
        """RigidBody3DState::serialize of the simulation's state, after replacing ( q, v ) and / or running updateMandMinv."""
        f = self.lib.ref_rb3d_sim_serialize_state
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        if q is not None:
            q, v = f64(q), f64(v)
        need = int(f(self.h, 1 if q is not None else 0, vp(q), vp(v), 1 if update else 0, None, 0))
        buf = np.zeros(need, dtype=np.uint8)
        assert int(f(self.h, 0, None, None, 0, vp(buf), need)) == need
        return buf.tobytes()

    @classmethod
    def from_snapshot(cls, blob, n):
        self = cls.__new__(cls)
        self.lib = lib = _lib("libref_rb3d.so")
        V = C.c_void_p
        lib.ref_rb3d_sim_from_snapshot.restype = V
        lib.ref_rb3d_sim_from_snapshot.argtypes = [V, C.c_uint64]
        lib.ref_rb3d_sim_destroy.argtypes = [V]
        lib.ref_rb3d_sim_active_set.restype = C.c_uint64
        lib.ref_rb3d_sim_active_set.argtypes = [V, V, V, C.c_uint64, V, V, V, V, V, V, V]
        raw = np.frombuffer(blob, dtype=np.uint8).copy()
        self.n, self.meshes = n, []
        self.h = lib.ref_rb3d_sim_from_snapshot(vp(raw), raw.shape[0])
        return self

    def active_set(self, q0, q1, cap=None):
        q0, q1 = f64(q0), f64(q1)
        cap = cap or 64 * self.n + 4096
        out = {"type": np.zeros(cap, np.uint32), "i": np.zeros(cap, np.uint32), "j": np.zeros(cap, np.uint32), "static": np.zeros(cap, np.uint32),
               "n": np.zeros((cap, 3)), "p": np.zeros((cap, 3)), "depth": np.zeros(cap)}
        na = int(self.lib.ref_rb3d_sim_active_set(self.h, vp(q0), vp(q1), cap, vp(out["type"]), vp(out["i"]), vp(out["j"]), vp(out["static"]), vp(out["n"]), vp(out["p"]), vp(out["depth"])))
        if na > cap:
            return self.active_set(q0, q1, cap=na)
        return {k: v[:na] for k, v in out.items()}
