"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

_lib = None


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h")) or f == "Makefile"]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B", "liboracle.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(LIB_PATH)
    vp, dp, up = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    sig = {
        "orc_ccd_coeffs": (None, [vp, vp, C.c_double, vp, vp, C.c_double, vp]),
        "orc_ccd_happens": (C.c_int, [vp, dp]),
        "orc_aabb_overlaps": (vp, [C.c_int, C.c_uint32, vp, C.c_int]),
        "orc_pairs_count": (C.c_uint64, [vp]),
        "orc_pairs_seconds": (C.c_double, [vp]),
        "orc_pairs_copy": (None, [vp, vp]),
        "orc_pairs_free": (None, [vp]),
        "orc_ball2d_create": (vp, [C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp]),
        "orc_ball2d_destroy": (None, [vp]),
        "orc_ball2d_flow": (None, [vp, C.c_int, vp, vp, C.c_double, vp, vp]),
        "orc_ball2d_active_set": (None, [vp, vp, vp, C.c_int]),
        "orc_ball2d_num_candidates": (C.c_uint64, [vp]),
        "orc_ball2d_num_active": (C.c_uint64, [vp]),
        "orc_ball2d_seconds_flow": (C.c_double, [vp]),
        "orc_ball2d_seconds_active": (C.c_double, [vp]),
        "orc_ball2d_copy_candidates": (None, [vp, vp]),
        "orc_ball2d_copy_active": (None, [vp, vp, vp, vp, vp, vp, vp]),
        "orc_ball2d_parallel_step": (C.c_double, [vp, C.c_int, vp, vp, C.c_double, vp, vp, C.c_int, vp]),
        "orc_ball2d_parallel_copy": (None, [vp, vp, vp]),
        "orc_ball2d_set_portals": (None, [vp, C.c_uint32, vp, vp, vp, vp, vp, vp]),
        "orc_ball2d_update_portals": (None, [vp, C.c_double, vp]),
        "orc_ball2d_enforce_portals": (None, [vp, vp, vp]),
        "orc_ball2d_portal_probe": (C.c_uint32, [vp, C.c_uint32, vp, C.c_double, vp]),
        "orc_ball2d_active_set_portals": (C.c_int, [vp, vp, vp, C.c_int]),
        "orc_ball2d_portals_num_regular": (C.c_uint64, [vp]),
        "orc_ball2d_portals_num_boxes": (C.c_uint64, [vp]),
        "orc_ball2d_portals_num_teleported": (C.c_uint64, [vp]),
        "orc_ball2d_portals_copy_boxes": (None, [vp, vp, vp]),
        "orc_ball2d_portals_copy_teleported": (None, [vp, vp, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ccd(q0a, q1a, ra, q0b, q1b, rb):
    lib = load()
    c = np.zeros(3)
    lib.orc_ccd_coeffs(_p(_f64(q0a)), _p(_f64(q1a)), float(ra), _p(_f64(q0b)), _p(_f64(q1b)), float(rb), _p(c))
    t = C.c_double()
    hit = lib.orc_ccd_happens(_p(c), C.byref(t))
    return c, bool(hit), t.value


def aabb_overlaps(aabbs, method="grid"):
    """(P,2) uint32 ascending pairs; method 'grid' = literal std::map/std::set algorithm, 'allpairs' = brute force."""
    lib = load()
    a = _f64(aabbs)
    h = lib.orc_aabb_overlaps(a.shape[1] // 2, a.shape[0], _p(a), 0 if method == "grid" else 1)
    n = lib.orc_pairs_count(h)
    out = np.zeros((n, 2), dtype=np.uint32)
    if n:
        lib.orc_pairs_copy(h, _p(out))
    secs = lib.orc_pairs_seconds(h)
    lib.orc_pairs_free(h)
    return out, secs


class Ball2DOracle:
    def __init__(self, scene):
        self.lib = load()
        self.n = scene["r"].shape[0]
        a = lambda k: _f64(scene[k])
        self._keep = [a("r"), a("m"), a("g"), a("plane_x"), a("plane_n"), a("drum_x"), a("drum_r")]
        r, m, g, px, pn, dx, dr = self._keep
        self.h = self.lib.orc_ball2d_create(self.n, _p(r), _p(m), _p(g), px.shape[0], _p(px), _p(pn), dx.shape[0], _p(dx), _p(dr))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_ball2d_destroy(self.h)
            self.h = None

    def flow(self, kind, q0, v0, dt):
        q0, v0 = _f64(q0), _f64(v0)
        q1, v1 = np.empty_like(q0), np.empty_like(v0)
        self.lib.orc_ball2d_flow(self.h, int(kind), _p(q0), _p(v0), float(dt), _p(q1), _p(v1))
        return q1, v1

    def active_set(self, q0, q1, method="grid"):
        q0, q1 = _f64(q0), _f64(q1)
        self.lib.orc_ball2d_active_set(self.h, _p(q0), _p(q1), 0 if method == "grid" else 1)
        nc = self.lib.orc_ball2d_num_candidates(self.h)
        na = self.lib.orc_ball2d_num_active(self.h)
        cand = np.zeros((nc, 2), dtype=np.uint32)
        if nc:
            self.lib.orc_ball2d_copy_candidates(self.h, _p(cand))
        out = {"type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32),
               "n": np.zeros((na, 2)), "p": np.zeros((na, 2)), "depth": np.zeros(na), "candidates": cand}
        if na:
            self.lib.orc_ball2d_copy_active(self.h, _p(out["type"]), _p(out["i"]), _p(out["j"]), _p(out["n"]), _p(out["p"]), _p(out["depth"]))
        out["seconds"] = self.lib.orc_ball2d_seconds_active(self.h)
        out["seconds_flow"] = self.lib.orc_ball2d_seconds_flow(self.h)
        return out

    # ---- N, Q, contact bases, constraint cache for the last active set (oracle/assembly2d.h) ----
    def assemble(self):
        sizes = np.zeros(3, dtype=np.uint64)
        self.lib.orc_ball2d_assemble.restype = C.c_int
        ok = self.lib.orc_ball2d_assemble(self.h, _p(sizes))
        nc, nn, nq = [int(x) for x in sizes]
        out = {"supported": bool(ok), "n_outer": np.zeros(nc + 1, np.int32), "n_inner": np.zeros(nn, np.int32), "n_values": np.zeros(nn), "q_outer": np.zeros(nc + 1, np.int32),
               "q_inner": np.zeros(nq, np.int32), "q_values": np.zeros(nq), "bases": np.zeros(4 * nc)}
        self.lib.orc_ball2d_copy_assembly(self.h, *[_p(out[k]) for k in ("n_outer", "n_inner", "n_values", "q_outer", "q_inner", "q_values", "bases")])
        return out

    def cache_store(self, r, ncomp):
        r = _f64(r)
        self.lib.orc_ball2d_cache_store(self.h, int(ncomp), _p(r))

    def cache_lookup(self, n_active, ncomp):
        out = np.zeros(max(1, n_active * ncomp))
        self.lib.orc_ball2d_cache_lookup.restype = C.c_uint64
        hits = int(self.lib.orc_ball2d_cache_lookup(self.h, int(ncomp), _p(out)))
        return out[: n_active * ncomp], hits

    # ---- multi-core step (oracle/ball2d_parallel.h): NOT reference behaviour, the optional second CPU baseline ----
    def parallel_step(self, kind, q0, v0, dt, keep_lists=True):
        """flow + active set on all host threads. Returns dict(q1, v1, seconds, n_candidates, n_active, n_static, threads[, candidates, active])."""
        q0, v0 = _f64(q0), _f64(v0)
        q1, v1 = np.empty_like(q0), np.empty_like(v0)
        out = np.zeros(4, dtype=np.uint64)
        secs = self.lib.orc_ball2d_parallel_step(self.h, int(kind), _p(q0), _p(v0), float(dt), _p(q1), _p(v1), 1 if keep_lists else 0, _p(out))
        res = {"q1": q1, "v1": v1, "seconds": float(secs), "n_candidates": int(out[0]), "n_active": int(out[1]), "n_static": int(out[2]), "threads": int(out[3])}
        if keep_lists:
            cand, act = np.zeros((int(out[0]), 2), np.uint32), np.zeros((int(out[1]), 2), np.uint32)
            self.lib.orc_ball2d_parallel_copy(self.h, _p(cand), _p(act))
            res["candidates"], res["active"] = cand, act
        return res

    # ---- portals (oracle/ball2d_portals.h) ----
    def set_portals(self, portals):
        """portals: dict with plane_a_x, plane_a_n, plane_b_x, plane_b_n (P,2), v (P), bounds (P)."""
        a = [_f64(portals[k]) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n", "v", "bounds")]
        self.nportals = a[4].shape[0]
        self.lib.orc_ball2d_set_portals(self.h, self.nportals, *[_p(x) for x in a])

    def update_portals(self, t):
        dx = np.zeros(self.nportals)
        self.lib.orc_ball2d_update_portals(self.h, float(t), _p(dx))
        return dx

    def enforce_portals(self, q, v):
        q, v = _f64(q).copy(), _f64(v).copy()
        self.lib.orc_ball2d_enforce_portals(self.h, _p(q), _p(v))
        return q, v

    def portal_probe(self, p, x, r):
        out = np.zeros(12)
        flags = self.lib.orc_ball2d_portal_probe(self.h, int(p), _p(_f64(x)), float(r), _p(out))
        return int(flags), out

    def active_set_portals(self, q0, q1, method="grid"):
        """None where the reference exits (a ball touching both planes of one portal)."""
        q0, q1 = _f64(q0), _f64(q1)
        if self.lib.orc_ball2d_active_set_portals(self.h, _p(q0), _p(q1), 0 if method == "grid" else 1) != 0:
            return None
        nc = self.lib.orc_ball2d_num_candidates(self.h)
        na = self.lib.orc_ball2d_num_active(self.h)
        cand = np.zeros((nc, 2), dtype=np.uint32)
        if nc:
            self.lib.orc_ball2d_copy_candidates(self.h, _p(cand))
        out = {"type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32),
               "n": np.zeros((na, 2)), "p": np.zeros((na, 2)), "depth": np.zeros(na), "candidates": cand}
        if na:
            self.lib.orc_ball2d_copy_active(self.h, _p(out["type"]), _p(out["i"]), _p(out["j"]), _p(out["n"]), _p(out["p"]), _p(out["depth"]))
        nbx = self.lib.orc_ball2d_portals_num_boxes(self.h)
        nt = self.lib.orc_ball2d_portals_num_teleported(self.h)
        out["n_regular"] = int(self.lib.orc_ball2d_portals_num_regular(self.h))
        out["box_body"], out["box_portal"] = np.zeros(nbx, np.uint32), np.zeros(nbx, np.uint32)
        if nbx:
            self.lib.orc_ball2d_portals_copy_boxes(self.h, _p(out["box_body"]), _p(out["box_portal"]))
        out["portal0"], out["portal1"] = np.zeros(nt, np.uint32), np.zeros(nt, np.uint32)
        out["x0"], out["x1"], out["kick"] = np.zeros((nt, 2)), np.zeros((nt, 2)), np.zeros((nt, 2))
        if nt:
            self.lib.orc_ball2d_portals_copy_teleported(self.h, _p(out["portal0"]), _p(out["portal1"]), _p(out["x0"]), _p(out["x1"]), _p(out["kick"]))
        out["seconds"] = self.lib.orc_ball2d_seconds_active(self.h)
        return out


class RB3DOracle:
    def __init__(self, scene):
        self.lib = lib = load()
        vp = C.c_void_p
        if not hasattr(lib, "_rb3d_bound"):
            lib.orc_rb3d_create.restype = vp
            lib.orc_rb3d_create.argtypes = [C.c_uint32, vp, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, vp, C.c_uint32, vp, vp]
            lib.orc_rb3d_destroy.argtypes = [vp]
            lib.orc_rb3d_add_mesh.restype = C.c_uint32
            lib.orc_rb3d_add_mesh.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp, vp]
            lib.orc_rb3d_flow.argtypes = [vp, C.c_int, vp, vp, C.c_double, vp, vp]
            lib.orc_rb3d_active_set.restype = C.c_int
            lib.orc_rb3d_active_set.argtypes = [vp, vp, vp, C.c_int]
            for f in ("orc_rb3d_num_candidates", "orc_rb3d_num_active"):
                getattr(lib, f).restype = C.c_uint64
                getattr(lib, f).argtypes = [vp]
            for f in ("orc_rb3d_seconds_flow", "orc_rb3d_seconds_active"):
                getattr(lib, f).restype = C.c_double
                getattr(lib, f).argtypes = [vp]
            lib.orc_rb3d_copy_candidates.argtypes = [vp, vp]
            lib.orc_rb3d_copy_active.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
            lib._rb3d_bound = True
        s = scene
        self.n = s["geo_of_body"].shape[0]
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        self._keep = [u32(s["geo_of_body"]), np.ascontiguousarray(s["fixed"], dtype=np.uint8), _f64(s["m"]), _f64(s["I0"]), _f64(s["g"]),
                      u32(s["geo_type"]), _f64(s["geo_r"]), _f64(s["geo_half"]), u32(s["geo_mesh"]), _f64(s["plane_x"]), _f64(s["plane_n"])]
        k = self._keep
        self.h = lib.orc_rb3d_create(self.n, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), k[5].shape[0], _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]),
                                     k[9].shape[0], _p(k[9]), _p(k[10]))
        if "cyl_r" in s and len(s["cyl_r"]):
            cx, ca, cr = _f64(s["cyl_x"]), _f64(s["cyl_axis"]), _f64(s["cyl_r"])
            lib.orc_rb3d_set_cylinders.argtypes = [vp, C.c_uint32, vp, vp, vp]
            lib.orc_rb3d_set_cylinders(self.h, cr.shape[0], _p(cx), _p(ca), _p(cr))
        for mesh in s["meshes"]:
            v, sm, hl = _f64(mesh["verts"]), _f64(mesh["samples"]), _f64(mesh["hull"])
            lib.orc_rb3d_add_mesh(self.h, v.shape[0], _p(v), sm.shape[0], _p(sm), hl.shape[0], _p(hl), _p(_f64(mesh["cell_delta"])), _p(u32(mesh["dims"])),
                                  _p(_f64(mesh["origin"])), _p(_f64(mesh["sdf"])))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_rb3d_destroy(self.h)
            self.h = None

    def flow(self, kind, q0, v0, dt, m_updated=False):
        """m_updated: M as RigidBody3DState::updateMandMinv leaves it (every flow after a simulation's first), see oracle/rb3d.h."""
        q0, v0 = _f64(q0), _f64(v0)
        q1, v1 = np.empty_like(q0), np.empty_like(v0)
        fn = self.lib.orc_rb3d_flow_m_updated if m_updated else self.lib.orc_rb3d_flow
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        fn.restype = None
        fn(self.h, int(kind), _p(q0), _p(v0), float(dt), _p(q1), _p(v1))
        return q1, v1

    def update_m_and_minv(self, q):
        """RigidBody3DState::updateMandMinv: ( I blocks, Iinv blocks ), 9 doubles per body, column-major as in M's value array."""
        q = _f64(q)
        I, Ii = np.zeros(9 * self.n), np.zeros(9 * self.n)
        self.lib.orc_rb3d_update_m_minv.argtypes = [C.c_void_p] * 4
        self.lib.orc_rb3d_update_m_minv.restype = None
        self.lib.orc_rb3d_update_m_minv(self.h, _p(q), _p(I), _p(Ii))
        return I, Ii

    def active_set(self, q0, q1, method="grid"):
        q0, q1 = _f64(q0), _f64(q1)
        ok = self.lib.orc_rb3d_active_set(self.h, _p(q0), _p(q1), 0 if method == "grid" else 1)
        nc = self.lib.orc_rb3d_num_candidates(self.h)
        na = self.lib.orc_rb3d_num_active(self.h)
        cand = np.zeros((nc, 2), dtype=np.uint32)
        if nc:
            self.lib.orc_rb3d_copy_candidates(self.h, _p(cand))
        out = {"type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32), "aux": np.zeros(na, np.uint32),
               "n": np.zeros((na, 3)), "p": np.zeros((na, 3)), "depth": np.zeros(na), "candidates": cand, "supported": bool(ok)}
        if na:
            self.lib.orc_rb3d_copy_active(self.h, _p(out["type"]), _p(out["i"]), _p(out["j"]), _p(out["aux"]), _p(out["n"]), _p(out["p"]), _p(out["depth"]))
        out["seconds"] = self.lib.orc_rb3d_seconds_active(self.h)
        out["seconds_flow"] = self.lib.orc_rb3d_seconds_flow(self.h)
        return out

    # ---- portals (oracle/rb3d_portals.h) ----
    def _bind_portals(self):
        lib, vp = self.lib, C.c_void_p
        if hasattr(lib, "_rb3d_portals_bound"):
            return
        lib.orc_rb3d_plane_frame.argtypes = [vp, vp, vp]
        lib.orc_rb3d_set_portals.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp]
        lib.orc_rb3d_portal_probe.restype = C.c_uint32
        lib.orc_rb3d_portal_probe.argtypes = [vp, C.c_uint32, vp, vp, vp]
        lib.orc_rb3d_enforce_portals.argtypes = [vp, vp]
        lib.orc_rb3d_active_set_portals.restype = C.c_int
        lib.orc_rb3d_active_set_portals.argtypes = [vp, vp, vp, C.c_int]
        for f in ("orc_rb3d_portals_num_regular", "orc_rb3d_portals_num_boxes", "orc_rb3d_portals_num_teleported"):
            getattr(lib, f).restype = C.c_uint64
            getattr(lib, f).argtypes = [vp]
        lib.orc_rb3d_portals_copy_boxes.argtypes = [vp, vp, vp]
        lib.orc_rb3d_portals_copy_teleported.argtypes = [vp] * 5
        lib._rb3d_portals_bound = True

    def plane_frame(self, x, n):
        self._bind_portals()
        out = np.zeros(9)
        self.lib.orc_rb3d_plane_frame(_p(_f64(x)), _p(_f64(n)), _p(out))
        return out

    def set_portals(self, portals):
        """portals: dict with plane_a_x, plane_a_n, plane_b_x, plane_b_n (P,3) and mult (P,3) int32."""
        self._bind_portals()
        a = [_f64(portals[k]) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n")]
        m = np.ascontiguousarray(portals["mult"], dtype=np.int32)
        self.nportals = m.shape[0]
        self.lib.orc_rb3d_set_portals(self.h, self.nportals, *[_p(x) for x in a], _p(m))

    def portal_probe(self, p, box, x):
        out = np.zeros(9)
        code = self.lib.orc_rb3d_portal_probe(self.h, int(p), _p(_f64(box)), _p(_f64(x)), _p(out))
        return int(code), out

    def enforce_portals(self, q):
        q = _f64(q).copy()
        self.lib.orc_rb3d_enforce_portals(self.h, _p(q))
        return q

    def active_set_portals(self, q0, q1, method="grid"):
        q0, q1 = _f64(q0), _f64(q1)
        ok = self.lib.orc_rb3d_active_set_portals(self.h, _p(q0), _p(q1), 0 if method == "grid" else 1)
        nc, na = self.lib.orc_rb3d_num_candidates(self.h), self.lib.orc_rb3d_num_active(self.h)
        cand = np.zeros((nc, 2), dtype=np.uint32)
        if nc:
            self.lib.orc_rb3d_copy_candidates(self.h, _p(cand))
        out = {"type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32), "aux": np.zeros(na, np.uint32),
               "n": np.zeros((na, 3)), "p": np.zeros((na, 3)), "depth": np.zeros(na), "candidates": cand, "supported": bool(ok)}
        if na:
            self.lib.orc_rb3d_copy_active(self.h, _p(out["type"]), _p(out["i"]), _p(out["j"]), _p(out["aux"]), _p(out["n"]), _p(out["p"]), _p(out["depth"]))
        nbx, nt = self.lib.orc_rb3d_portals_num_boxes(self.h), self.lib.orc_rb3d_portals_num_teleported(self.h)
        out["n_regular"] = int(self.lib.orc_rb3d_portals_num_regular(self.h))
        out["box_body"], out["box_portal"] = np.zeros(nbx, np.uint32), np.zeros(nbx, np.uint32)
        if nbx:
            self.lib.orc_rb3d_portals_copy_boxes(self.h, _p(out["box_body"]), _p(out["box_portal"]))
        out["portal0"], out["portal1"], out["x0"], out["x1"] = np.zeros(nt, np.uint32), np.zeros(nt, np.uint32), np.zeros((nt, 3)), np.zeros((nt, 3))
        if nt:
            self.lib.orc_rb3d_portals_copy_teleported(self.h, _p(out["portal0"]), _p(out["portal1"]), _p(out["x0"]), _p(out["x1"]))
        return out


class RB2DOracle:
    def __init__(self, scene):
        self.lib = lib = load()
        vp = C.c_void_p
        if not hasattr(lib, "_rb2d_bound"):
            lib.orc_rb2d_create.restype = vp
            lib.orc_rb2d_create.argtypes = [C.c_uint32, vp, vp, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint32, vp, vp]
            lib.orc_rb2d_destroy.argtypes = [vp]
            lib.orc_rb2d_flow.argtypes = [vp, C.c_int, vp, vp, C.c_double, vp, vp]
            lib.orc_rb2d_active_set.restype = C.c_int
            lib.orc_rb2d_active_set.argtypes = [vp, vp, vp, C.c_int]
            for f in ("orc_rb2d_num_candidates", "orc_rb2d_num_active"):
                getattr(lib, f).restype = C.c_uint64
                getattr(lib, f).argtypes = [vp]
            lib.orc_rb2d_copy_candidates.argtypes = [vp, vp]
            lib.orc_rb2d_copy_active.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
            lib.orc_rb2d_set_portals.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp, vp]
            lib.orc_rb2d_update_portals.argtypes = [vp, C.c_double, vp]
            lib.orc_rb2d_enforce_portals.argtypes = [vp, vp, vp]
            lib.orc_rb2d_portal_probe.restype = C.c_uint32
            lib.orc_rb2d_portal_probe.argtypes = [vp, C.c_uint32, vp, vp, vp]
            lib.orc_rb2d_active_set_portals.restype = C.c_int
            lib.orc_rb2d_active_set_portals.argtypes = [vp, vp, vp, C.c_int]
            for f in ("orc_rb2d_portals_num_regular", "orc_rb2d_portals_num_boxes", "orc_rb2d_portals_num_teleported"):
                getattr(lib, f).restype = C.c_uint64
                getattr(lib, f).argtypes = [vp]
            lib.orc_rb2d_portals_copy_boxes.argtypes = [vp, vp, vp]
            lib.orc_rb2d_portals_copy_teleported.argtypes = [vp] * 8
            lib._rb2d_bound = True
        s = scene
        self.n = s["geo_of_body"].shape[0]
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        k = self._keep = [u32(s["geo_of_body"]), np.ascontiguousarray(s["fixed"], dtype=np.uint8), _f64(s["M"]), _f64(s["g"]), u32(s["geo_type"]),
                          _f64(s["geo_r"]), _f64(s["geo_half"]), _f64(s["plane_x"]), _f64(s["plane_n"])]
        self.h = lib.orc_rb2d_create(self.n, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), k[4].shape[0], _p(k[4]), _p(k[5]), _p(k[6]), k[7].shape[0], _p(k[7]), _p(k[8]))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_rb2d_destroy(self.h)
            self.h = None

    def flow(self, kind, q0, v0, dt):
        q0, v0 = _f64(q0), _f64(v0)
        q1, v1 = np.empty_like(q0), np.empty_like(v0)
        self.lib.orc_rb2d_flow(self.h, int(kind), _p(q0), _p(v0), float(dt), _p(q1), _p(v1))
        return q1, v1

    def active_set(self, q0, q1, method="grid"):
        q0, q1 = _f64(q0), _f64(q1)
        ok = self.lib.orc_rb2d_active_set(self.h, _p(q0), _p(q1), 0 if method == "grid" else 1)
        nc, na = self.lib.orc_rb2d_num_candidates(self.h), self.lib.orc_rb2d_num_active(self.h)
        cand = np.zeros((nc, 2), dtype=np.uint32)
        if nc:
            self.lib.orc_rb2d_copy_candidates(self.h, _p(cand))
        out = {"type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32), "aux": np.zeros(na, np.uint32),
               "n": np.zeros((na, 2)), "p": np.zeros((na, 2)), "depth": np.zeros(na), "candidates": cand, "supported": bool(ok)}
        if na:
            self.lib.orc_rb2d_copy_active(self.h, _p(out["type"]), _p(out["i"]), _p(out["j"]), _p(out["aux"]), _p(out["n"]), _p(out["p"]), _p(out["depth"]))
        return out

    # ---- portals (oracle/rb2d_portals.h) ----
    def set_portals(self, portals):
        a = [_f64(portals[k]) for k in ("plane_a_x", "plane_a_n", "plane_b_x", "plane_b_n", "v", "bounds")]
        self.nportals = a[4].shape[0]
        self.lib.orc_rb2d_set_portals(self.h, self.nportals, *[_p(x) for x in a])

    def update_portals(self, t):
        dx = np.zeros(self.nportals)
        self.lib.orc_rb2d_update_portals(self.h, float(t), _p(dx))
        return dx

    def enforce_portals(self, q, v):
        q, v = _f64(q).copy(), _f64(v).copy()
        self.lib.orc_rb2d_enforce_portals(self.h, _p(q), _p(v))
        return q, v

    def portal_probe(self, p, box, x):
        out = np.zeros(4)
        touch = self.lib.orc_rb2d_portal_probe(self.h, int(p), _p(_f64(box)), _p(_f64(x)), _p(out))
        return int(touch), out

    def active_set_portals(self, q0, q1, method="grid"):
        q0, q1 = _f64(q0), _f64(q1)
        ok = self.lib.orc_rb2d_active_set_portals(self.h, _p(q0), _p(q1), 0 if method == "grid" else 1)
        nc, na = self.lib.orc_rb2d_num_candidates(self.h), self.lib.orc_rb2d_num_active(self.h)
        cand = np.zeros((nc, 2), dtype=np.uint32)
        if nc:
            self.lib.orc_rb2d_copy_candidates(self.h, _p(cand))
        out = {"type": np.zeros(na, np.uint32), "i": np.zeros(na, np.uint32), "j": np.zeros(na, np.uint32), "aux": np.zeros(na, np.uint32),
               "n": np.zeros((na, 2)), "p": np.zeros((na, 2)), "depth": np.zeros(na), "candidates": cand, "supported": bool(ok)}
        if na:
            self.lib.orc_rb2d_copy_active(self.h, _p(out["type"]), _p(out["i"]), _p(out["j"]), _p(out["aux"]), _p(out["n"]), _p(out["p"]), _p(out["depth"]))
        nbx, nt = self.lib.orc_rb2d_portals_num_boxes(self.h), self.lib.orc_rb2d_portals_num_teleported(self.h)
        out["n_regular"] = int(self.lib.orc_rb2d_portals_num_regular(self.h))
        out["box_body"], out["box_portal"] = np.zeros(nbx, np.uint32), np.zeros(nbx, np.uint32)
        if nbx:
            self.lib.orc_rb2d_portals_copy_boxes(self.h, _p(out["box_body"]), _p(out["box_portal"]))
        out["portal0"], out["portal1"] = np.zeros(nt, np.uint32), np.zeros(nt, np.uint32)
        for k in ("x0", "x1", "delta0", "delta1", "kick"):
            out[k] = np.zeros((nt, 2))
        if nt:
            self.lib.orc_rb2d_portals_copy_teleported(self.h, _p(out["portal0"]), _p(out["portal1"]), _p(out["x0"]), _p(out["x1"]), _p(out["delta0"]), _p(out["delta1"]), _p(out["kick"]))
        return out
