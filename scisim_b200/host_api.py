"""Host-side mirror (Python) of the reference interfaces on the hot path, over the C ABI.

Names, argument meaning and error behaviour follow the reference:
  UnconstrainedMap::flow( q0, v0, fsys, iteration, dt, q1, v1 )       scisim/UnconstrainedMaps/UnconstrainedMap.h:33
  ConstrainedSystem::computeActiveSet( q0, qp, v, active_set )        scisim/Constraints/ConstrainedSystem.h:20
  SpatialGridDetector::getPotentialOverlaps( aabbs, overlaps )        ball2d/SpatialGridDetector.h:39
Every compute call goes to the CUDA library; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import SG_MAP_DMV, SG_MAP_M_UPDATED, SG_MAP_SPLIT_HAM, SG_MAP_SYMPLECTIC_EULER, SG_MAP_VERLET, SG_OUT_ALL, SgContacts, SgPairs, SciSimB200Error


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One sg_ctx: one GPU, one stream, library-owned device buffers."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.sg_create(C.byref(h), int(device))
        if rc != 0:
            raise SciSimB200Error("sg_create failed (%d): %s" % (rc, self.lib.sg_last_error(None).decode()))
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise SciSimB200Error("libscisim_b200 error %d: %s" % (rc, self.lib.sg_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.check(self.lib.sg_synchronize(self.h))

    def stream(self):
        return self.lib.sg_stream(self.h)

    def launch_count(self):
        return int(self.lib.sg_launch_count(self.h))

    def timer_begin(self):
        self.check(self.lib.sg_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_double()
        self.check(self.lib.sg_timer_end(self.h, C.byref(ms)))
        return float(ms.value)

    def flush_l2(self):
        self.check(self.lib.sg_flush_l2(self.h))

    def profile_enable(self, on=True):
        self.check(self.lib.sg_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self.check(self.lib.sg_profile_reset(self.h))

    def profile(self):
        """{kernel name: (launches, device ms, algorithmic bytes)} accumulated since the last reset."""
        self.synchronize()
        out = {}
        for k in range(self.lib.sg_profile_count(self.h)):
            name = C.c_char_p()
            n = C.c_uint64()
            ms = C.c_double()
            by = C.c_double()
            self.check(self.lib.sg_profile_get(self.h, k, C.byref(name), C.byref(n), C.byref(ms), C.byref(by)))
            out[name.value.decode()] = (int(n.value), float(ms.value), float(by.value))
        return out

    def pinned(self, shape, dtype=np.float64):
        """A numpy array over pinned host memory owned by this context (freed with the context's process)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self.lib.sg_host_alloc(self.h, max(n, 1), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class SpatialGridDetector:
    """Drop-in for the reference's namespace SpatialGridDetector (rigidbody2d: SpatialGrid)."""

    @staticmethod
    def getPotentialOverlaps(aabbs, ctx=None):
        """aabbs: (n, 2*dim) rows [lo, hi]. Returns a (P, 2) uint32 array of (i<j), ascending -- the
        iteration order of the std::set the reference fills."""
        ctx = ctx or default_context()
        a = _f64(aabbs)
        if a.ndim != 2 or a.shape[1] not in (4, 6):
            raise ValueError("aabbs must be (n,4) or (n,6)")
        out = SgPairs()
        ctx.check(ctx.lib.sg_candidate_pairs(ctx.h, a.shape[1] // 2, a.shape[0], _ptr(a), C.byref(out)))
        if out.n == 0:
            return np.zeros((0, 2), dtype=np.uint32)
        return np.ctypeslib.as_array(out.ij, shape=(int(out.n), 2)).copy()


class ActiveSet:
    """The active set as plain arrays in the reference's active_set order (one row per Constraint)."""

    def __init__(self, c, copy=True):
        na = int(c.n_active)
        self.dim = int(c.dim)
        self.n_candidates = int(c.n_candidates)
        self.n_active = na
        self.n_body_body = int(c.n_body_body)
        self.n_drum = int(c.n_drum)
        self.n_plane = int(c.n_plane)

        def arr(p, shape, dt):
            if not p or na == 0:
                return None if not p else np.zeros(shape, dtype=dt)
            a = np.ctypeslib.as_array(p, shape=shape)
            return a.copy() if copy else a

        self.type = arr(c.type, (na,), np.uint32)
        self.i = arr(c.i, (na,), np.uint32)
        self.j = arr(c.j, (na,), np.uint32)
        self.n = arr(c.n, (na, self.dim), np.float64)
        self.p = arr(c.p, (na, self.dim), np.float64)
        self.depth = arr(c.depth, (na,), np.float64)
        self.aux = arr(c.aux, (na,), np.uint32)
        if c.cand_ij and self.n_candidates > 0:
            a = np.ctypeslib.as_array(c.cand_ij, shape=(self.n_candidates, 2))
            self.candidates = a.copy() if copy else a
        elif c.cand_ij:
            self.candidates = np.zeros((0, 2), dtype=np.uint32)
        else:
            self.candidates = None


class PlanarPortal:
    """ball2d/Portals/PlanarPortal.h: PlanarPortal( plane_a, plane_b, velocity, bounds ); planes as (point, normal)."""

    def __init__(self, plane_a_x, plane_a_n, plane_b_x, plane_b_n, velocity=0.0, bounds=0.0):
        self.plane_a_x, self.plane_a_n = _f64(plane_a_x), _f64(plane_a_n)
        self.plane_b_x, self.plane_b_n = _f64(plane_b_x), _f64(plane_b_n)
        self.velocity, self.bounds = float(velocity), float(bounds)

    def isLeesEdwards(self):
        return self.velocity != 0.0

    @staticmethod
    def from_arrays(portals):
        """The dict layout of scenes.ball2d_periodic -> list of PlanarPortal."""
        return [PlanarPortal(portals["plane_a_x"][p], portals["plane_a_n"][p], portals["plane_b_x"][p], portals["plane_b_n"][p], portals["v"][p], portals["bounds"][p])
                for p in range(len(portals["v"]))]


class TeleportedInfo:
    """What the portal branch adds to an ActiveSet (sg_teleported, include/scisim_b200.h)."""

    def __init__(self, t, dim=2):
        nb, nt = int(t.n_boxes), int(t.n_teleported)
        self.n_regular = int(t.n_regular)
        self.n_teleported = nt
        arr = lambda p, shape, dt: np.ctypeslib.as_array(p, shape=shape).copy() if shape[0] > 0 else np.zeros(shape, dtype=dt)
        self.box_body = arr(t.box_body, (nb,), np.uint32)
        self.box_portal = arr(t.box_portal, (nb,), np.uint32)
        self.portal0 = arr(t.portal0, (nt,), np.uint32)
        self.portal1 = arr(t.portal1, (nt,), np.uint32)
        self.x0 = arr(t.x0, (nt, dim), np.float64)
        self.x1 = arr(t.x1, (nt, dim), np.float64)
        self.kick = arr(t.kick, (nt, 2), np.float64) if t.kick else None  # 2-D sims only (Lees-Edwards)
        # rigidbody2d only: TeleportedCircleCircleConstraint's displacements
        self.delta0 = arr(t.delta0, (nt, 2), np.float64) if t.delta0 else None
        self.delta1 = arr(t.delta1, (nt, 2), np.float64) if t.delta1 else None


class PlanarPortal3D:
    """rigidbody3d/Portals/PlanarPortal.h: PlanarPortal( plane_a, plane_b, portal_multiplier ); planes as (point, normal)."""

    def __init__(self, plane_a_x, plane_a_n, plane_b_x, plane_b_n, multiplier=(1, 1, 1)):
        self.plane_a_x, self.plane_a_n = _f64(plane_a_x), _f64(plane_a_n)
        self.plane_b_x, self.plane_b_n = _f64(plane_b_x), _f64(plane_b_n)
        self.multiplier = np.ascontiguousarray(multiplier, dtype=np.int32)

    @staticmethod
    def from_arrays(portals):
        """The dict layout of scenes.rb3d_periodic_spheres -> list of PlanarPortal3D."""
        return [PlanarPortal3D(portals["plane_a_x"][p], portals["plane_a_n"][p], portals["plane_b_x"][p], portals["plane_b_n"][p], portals["mult"][p])
                for p in range(len(portals["mult"]))]


class Ball2DState:
    """Static part of ball2d/Ball2DState.h: radii, per-ball masses, gravity, planes, drums, planar portals."""

    def __init__(self, r, m, g=(0.0, 0.0), plane_x=None, plane_n=None, drum_x=None, drum_r=None, planar_portals=None):
        self.r = _f64(r)
        self.m = _f64(m)
        assert self.r.shape == self.m.shape and self.r.ndim == 1
        self.g = _f64(g)
        self.plane_x = _f64(plane_x if plane_x is not None else np.zeros((0, 2)))
        self.plane_n = _f64(plane_n if plane_n is not None else np.zeros((0, 2)))
        self.drum_x = _f64(drum_x if drum_x is not None else np.zeros((0, 2)))
        self.drum_r = _f64(drum_r if drum_r is not None else np.zeros((0,)))
        self.planar_portals = list(planar_portals) if planar_portals is not None else []

    def nballs(self):
        return self.r.shape[0]

    def numPlanarPortals(self):
        return len(self.planar_portals)


class _Ball2DMap:
    kind = None
    _name = None

    def name(self):
        return self._name

    def flow(self, q0, v0, fsys, iteration, dt):
        """Same contract as UnconstrainedMap::flow; returns (q1, v1) instead of filling caller vectors."""
        assert iteration > 0
        return fsys._flow(self.kind, q0, v0, dt)


class SymplecticEulerMap(_Ball2DMap):
    """ball2d/SymplecticEulerMap.cpp"""
    kind = SG_MAP_SYMPLECTIC_EULER
    _name = "symplectic_euler"


class VerletMap(_Ball2DMap):
    """ball2d/VerletMap.cpp"""
    kind = SG_MAP_VERLET
    _name = "verlet"


class Ball2DSim:
    """GPU-backed FlowableSystem + ConstrainedSystem for ball2d (ball2d/Ball2DSim.h:26)."""

    def __init__(self, state, device=0, ctx=None):
        self.ctx = ctx or Context(device)
        self.state = state
        lib, h = self.ctx.lib, self.ctx.h
        n = state.nballs()
        self.ctx.check(lib.sg_ball2d_set_bodies(h, n, _ptr(state.r), _ptr(state.m)))
        self.ctx.check(lib.sg_ball2d_set_gravity(h, _ptr(state.g)))
        self.ctx.check(lib.sg_ball2d_set_planes(h, state.plane_x.shape[0], _ptr(state.plane_x), _ptr(state.plane_n)))
        self.ctx.check(lib.sg_ball2d_set_drums(h, state.drum_x.shape[0], _ptr(state.drum_x), _ptr(state.drum_r)))
        pp = state.planar_portals
        if pp:
            cat = lambda f: _f64(np.array([f(p) for p in pp], dtype=np.float64))
            arrs = [cat(lambda p: p.plane_a_x), cat(lambda p: p.plane_a_n), cat(lambda p: p.plane_b_x), cat(lambda p: p.plane_b_n),
                    cat(lambda p: p.velocity), cat(lambda p: p.bounds)]
            self.ctx.check(lib.sg_ball2d_set_portals(h, len(pp), *[_ptr(a) for a in arrs]))
        else:
            self.ctx.check(lib.sg_ball2d_set_portals(h, 0, None, None, None, None, None, None))

    def name(self):
        return "ball_2d"

    # ---- portals (ball2d/Ball2DSim.cpp:327-366) ----
    def updatePeriodicBoundaryConditionsStartOfStep(self, next_iteration, dt):
        """Moves the Lees-Edwards portals to t = next_iteration * dt; returns their tangential offsets."""
        dx = np.zeros(max(1, self.state.numPlanarPortals()))
        self.ctx.check(self.ctx.lib.sg_ball2d_update_portals(self.ctx.h, float(next_iteration * dt), _ptr(dx)))
        return dx[: self.state.numPlanarPortals()]

    def enforcePeriodicBoundaryConditions(self, q, v):
        """Teleports the balls that left through a portal (and adds the Lees-Edwards velocity); returns new (q, v)."""
        q = _f64(q).copy()
        v = _f64(v).copy()
        self.ctx.check(self.ctx.lib.sg_ball2d_enforce_portals(self.ctx.h, _ptr(q), _ptr(v)))
        return q, v

    def teleported(self):
        """Details of the last active set computed with portals (teleported boxes, constructor arguments of the teleported contacts)."""
        from ._lib import SgTeleported
        t = SgTeleported()
        self.ctx.check(self.ctx.lib.sg_ball2d_teleported(self.ctx.h, C.byref(t)))
        return TeleportedInfo(t)

    def nqdofs(self):
        return 2 * self.state.nballs()

    nvdofs = nqdofs

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        q0 = _f64(q0)
        v0 = _f64(v0)
        assert q0.size == self.nqdofs() and v0.size == self.nqdofs()
        q1 = np.empty_like(q0) if q1 is None else q1
        v1 = np.empty_like(v0) if v1 is None else v1
        self.ctx.check(self.ctx.lib.sg_ball2d_flow(self.ctx.h, kind, _ptr(q0), _ptr(v0), float(dt), _ptr(q1), _ptr(v1)))
        return q1, v1

    def computeActiveSet(self, q0, qp, v=None, flags=SG_OUT_ALL, copy=True, resident=False):
        """ConstrainedSystem::computeActiveSet( q0, qp, v, active_set ); v is unused by ball2d.
        resident=True: (q0, qp) are the input and output of the last flow() on this sim -- what ImpactMap::flow passes --
        and are not uploaded again (SG_IN_RESIDENT)."""
        from ._lib import SG_IN_RESIDENT
        q0 = _f64(q0)
        qp = _f64(qp)
        assert q0.size == self.nqdofs() and qp.size == self.nqdofs()
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_ball2d_active_set(self.ctx.h, _ptr(q0), _ptr(qp), int(flags) | (SG_IN_RESIDENT if resident else 0), C.byref(c)))
        return ActiveSet(c, copy=copy)

    # ---- device-side assembly for the solver's first step (ImpactMap.cpp:106-110, Ball2DSim.cpp:188-201) and the impulse cache ----
    def assemble(self, flags=7):
        """N (pruned, CSC), Q = N^T Minv N (CSC, sorted rows) and the contact bases of the LAST active set, computed on the device."""
        from ._lib import SgAssembly
        a = SgAssembly()
        self.ctx.check(self.ctx.lib.sg_ball2d_assemble(self.ctx.h, int(flags), C.byref(a)))
        nc, nn, nq = int(a.n_constraints), int(a.n_nnz), int(a.q_nnz)
        arr = lambda p, n, dt: (np.ctypeslib.as_array(p, shape=(n,)).copy() if (p and n > 0) else np.zeros(n, dtype=dt))
        return {"n_constraints": nc, "n_dofs": int(a.n_dofs),
                "n_outer": arr(a.n_outer, nc + 1, np.int32), "n_inner": arr(a.n_inner, nn, np.int32), "n_values": arr(a.n_values, nn, np.float64),
                "q_outer": arr(a.q_outer, nc + 1, np.int32), "q_inner": arr(a.q_inner, nq, np.int32), "q_values": arr(a.q_values, nq, np.float64),
                "bases": arr(a.bases, 4 * nc, np.float64)}

    def cacheConstraints(self, r, ncomp=1):
        """cacheConstraint for every constraint of the current active set (ball2d/ConstraintCache.cpp:20-60)."""
        r = _f64(r)
        self.ctx.check(self.ctx.lib.sg_ball2d_cache_store(self.ctx.h, int(ncomp), _ptr(r)))

    def getCachedConstraintImpulses(self, n_active, ncomp=1):
        """getCachedConstraintImpulse for every constraint of the current active set: ( values, constraints found )."""
        out = np.zeros(max(1, n_active * ncomp))
        hits = C.c_uint64()
        self.ctx.check(self.ctx.lib.sg_ball2d_cache_lookup(self.ctx.h, int(ncomp), _ptr(out), C.byref(hits)))
        return out[: n_active * ncomp], int(hits.value)

    def clearConstraintCache(self):
        self.ctx.check(self.ctx.lib.sg_ball2d_cache_clear(self.ctx.h))

    # ---- state I/O: Ball2DState's binary snapshot (ball2d/Ball2DState.cpp:259-312) ----
    def serializeState(self, which=1):
        """bytes of Ball2DState::serialize for the device-resident state: which = 0 the uploaded ( q0, v0 ), 1 the last flow's ( q1, v1 )."""
        need = C.c_uint64()
        self.ctx.check(self.ctx.lib.sg_ball2d_state_serialize(self.ctx.h, int(which), None, 0, C.byref(need)))
        buf = np.zeros(int(need.value), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sg_ball2d_state_serialize(self.ctx.h, int(which), _ptr(buf), buf.shape[0], C.byref(need)))
        return buf.tobytes()

    @staticmethod
    def deserializeState(blob, ctx):
        """Ball2DState::deserialize: a sim configured from a snapshot, its ( q, v ) uploaded."""
        sim = Ball2DSim.__new__(Ball2DSim)
        sim.ctx = ctx
        buf = np.frombuffer(blob, dtype=np.uint8).copy()
        ctx.check(ctx.lib.sg_ball2d_state_deserialize(ctx.h, _ptr(buf), buf.shape[0]))
        n = int(np.frombuffer(blob[:8], dtype=np.int64)[0]) // 2
        sim.state = Ball2DState(np.ones(n), np.ones(n))   # the context holds the real values; only nballs() is used on the host side
        return sim

    # ---- resident path (state stays in HBM) ----
    def upload(self, q, v):
        q = _f64(q)
        v = _f64(v)
        self.ctx.check(self.ctx.lib.sg_ball2d_upload(self.ctx.h, _ptr(q), _ptr(v)))

    def step(self, umap, dt):
        """flow(q0,v0)->(q1,v1) then the active set on (q0,q1), all on the device. Returns (P_c, P_a)."""
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_ball2d_step(self.ctx.h, umap.kind, float(dt), C.byref(c)))
        return int(c.n_candidates), int(c.n_active)

    def fetch(self, flags=SG_OUT_ALL, want_state=True):
        n = self.nqdofs()
        q1 = np.empty(n) if want_state else None
        v1 = np.empty(n) if want_state else None
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_ball2d_fetch(self.ctx.h, int(flags), _ptr(q1) if want_state else None, _ptr(v1) if want_state else None, C.byref(c)))
        return q1, v1, ActiveSet(c)


class MultiGpuBall2DSim:
    """Ball2DSim over several GPUs of ONE process (sg_multi, include/scisim_b200.h): the same calls, global vectors and
    indices; the scene is cut into x-slabs behind the interface, the lists come back merged in the reference's order."""

    def __init__(self, state, devices):
        if state.planar_portals:
            raise SciSimB200Error("portals are not supported in slab mode")
        self.lib = _lib.load()
        self.state = state
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self.lib.sg_create_multi(C.byref(h), len(devices), devs)
        if rc != 0:
            raise SciSimB200Error("sg_create_multi failed (%d): %s" % (rc, self.lib.sg_multi_last_error(None).decode()))
        self.h = h
        self.world = len(devices)
        n = state.nballs()
        self.check(self.lib.sg_multi_ball2d_set_bodies(h, n, _ptr(state.r), _ptr(state.m)))
        self.check(self.lib.sg_multi_ball2d_set_gravity(h, _ptr(state.g)))
        self.check(self.lib.sg_multi_ball2d_set_planes(h, state.plane_x.shape[0], _ptr(state.plane_x), _ptr(state.plane_n)))
        self.check(self.lib.sg_multi_ball2d_set_drums(h, state.drum_x.shape[0], _ptr(state.drum_x), _ptr(state.drum_r)))

    def check(self, rc):
        if rc != 0:
            raise SciSimB200Error("libscisim_b200 (multi) error %d: %s" % (rc, self.lib.sg_multi_last_error(self.h).decode()))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sg_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nqdofs(self):
        return 2 * self.state.nballs()

    nvdofs = nqdofs

    def set_rebalance(self, every_n_uploads=0, ghost_cap=0):
        self.check(self.lib.sg_multi_set_rebalance(self.h, int(every_n_uploads), int(ghost_cap)))

    def partition_info(self):
        cuts = np.zeros(self.world + 1)
        owned = np.zeros(self.world, dtype=np.uint32)
        ghosts = np.zeros(2 * self.world, dtype=np.uint32)
        npart = C.c_uint64()
        self.check(self.lib.sg_multi_partition_info(self.h, _ptr(cuts), _ptr(owned), _ptr(ghosts), C.byref(npart)))
        return {"cuts": cuts, "n_owned": owned, "ghosts": ghosts.reshape(-1, 2), "n_partitions": int(npart.value)}

    def slab_stats(self):
        """per slab: (ghosts on side 0, ghosts on side 1, halo packs that had to scan all bodies, candidate-list capacity)"""
        out = []
        for k in range(self.world):
            a = np.zeros(4, dtype=np.uint32)
            self.lib.sg_ball2d_slab_stats(self.lib.sg_multi_context(self.h, k), _ptr(a))
            out.append(tuple(int(x) for x in a))
        return out

    def slab_launch_counts(self):
        return [int(self.lib.sg_launch_count(self.lib.sg_multi_context(self.h, k))) for k in range(self.world)]

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        q0, v0 = _f64(q0), _f64(v0)
        assert q0.size == self.nqdofs() and v0.size == self.nqdofs()
        q1 = np.empty_like(q0) if q1 is None else q1
        v1 = np.empty_like(v0) if v1 is None else v1
        self.check(self.lib.sg_multi_ball2d_flow(self.h, kind, _ptr(q0), _ptr(v0), float(dt), _ptr(q1), _ptr(v1)))
        return q1, v1

    def computeActiveSet(self, q0, qp, v=None, flags=SG_OUT_ALL, copy=True, resident=False):
        from ._lib import SG_IN_RESIDENT
        q0, qp = _f64(q0), _f64(qp)
        c = SgContacts()
        self.check(self.lib.sg_multi_ball2d_active_set(self.h, _ptr(q0), _ptr(qp), int(flags) | (SG_IN_RESIDENT if resident else 0), C.byref(c)))
        return ActiveSet(c, copy=copy)

    def upload(self, q, v):
        q, v = _f64(q), _f64(v)
        self.check(self.lib.sg_multi_ball2d_upload(self.h, _ptr(q), _ptr(v)))

    def step(self, umap, dt):
        c = SgContacts()
        self.check(self.lib.sg_multi_ball2d_step(self.h, umap.kind, float(dt), C.byref(c)))
        return int(c.n_candidates), int(c.n_active)

    def fetch(self, flags=SG_OUT_ALL, want_state=True, copy=True):
        n = self.nqdofs()
        q1 = np.empty(n) if want_state else None
        v1 = np.empty(n) if want_state else None
        c = SgContacts()
        self.check(self.lib.sg_multi_ball2d_fetch(self.h, int(flags), _ptr(q1) if want_state else None, _ptr(v1) if want_state else None, C.byref(c)))
        return q1, v1, ActiveSet(c, copy=copy)


# ---- rigidbody3d -------------------------------------------------------------------------------------
GEO_BOX, GEO_SPHERE, GEO_MESH = 0, 1, 3


class TriangleMesh:
    """What RigidBodyTriangleMesh holds for the hot path (rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:54-102)."""

    def __init__(self, verts, samples, hull, cell_delta, dims, origin, sdf):
        self.verts, self.samples, self.hull = _f64(verts).reshape(-1, 3), _f64(samples).reshape(-1, 3), _f64(hull).reshape(-1, 3)
        self.cell_delta, self.origin = _f64(cell_delta), _f64(origin)
        self.dims = np.ascontiguousarray(dims, dtype=np.uint32)
        self.sdf = _f64(sdf).ravel()
        assert self.sdf.size == int(np.prod(self.dims.astype(np.int64)))


class RigidBody3DState:
    """Static part of rigidbody3d/RigidBody3DState.h: geometry list, per-body geometry index / fixed flag / mass /
    body-frame inertia, gravity, static planes, static cylinders."""

    def __init__(self, geo_type, geo_r, geo_half, geo_mesh, meshes, geo_of_body, fixed, m, I0, g=(0.0, 0.0, 0.0), plane_x=None, plane_n=None,
                 cyl_x=None, cyl_axis=None, cyl_r=None, planar_portals=None):
        self.planar_portals = list(planar_portals) if planar_portals is not None else []
        self.cyl_x = _f64(cyl_x if cyl_x is not None else np.zeros((0, 3))).reshape(-1, 3)
        self.cyl_axis = _f64(cyl_axis if cyl_axis is not None else np.zeros((0, 3))).reshape(-1, 3)
        self.cyl_r = _f64(cyl_r if cyl_r is not None else np.zeros(0))
        self.geo_type = np.ascontiguousarray(geo_type, dtype=np.uint32)
        self.geo_r = _f64(geo_r)
        self.geo_half = _f64(geo_half).reshape(-1, 3)
        self.geo_mesh = np.ascontiguousarray(geo_mesh, dtype=np.uint32)
        self.meshes = list(meshes)
        self.geo_of_body = np.ascontiguousarray(geo_of_body, dtype=np.uint32)
        self.fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
        self.m = _f64(m)
        self.I0 = _f64(I0).reshape(-1, 3)
        self.g = _f64(g)
        self.plane_x = _f64(plane_x if plane_x is not None else np.zeros((0, 3))).reshape(-1, 3)
        self.plane_n = _f64(plane_n if plane_n is not None else np.zeros((0, 3))).reshape(-1, 3)

    def nbodies(self):
        return self.geo_of_body.shape[0]


class _RB3DMap:
    kind = None
    _name = None

    def name(self):
        return self._name

    def flow(self, q0, v0, fsys, iteration, dt):
        assert iteration > 0
        return fsys._flow(self.kind, q0, v0, dt)


class SplitHamMap(_RB3DMap):
    """rigidbody3d/UnconstrainedMaps/SplitHamMap.cpp"""
    kind = SG_MAP_SPLIT_HAM
    _name = "split_ham"


class DMVMap(_RB3DMap):
    """rigidbody3d/UnconstrainedMaps/DMVMap.cpp"""
    kind = SG_MAP_DMV
    _name = "dmv"


class ExponentialEulerMap(_RB3DMap):
    """rigidbody3d/UnconstrainedMaps/ExponentialEulerMap.cpp"""
    kind = 4  # SG_MAP_EXPONENTIAL_EULER
    _name = "exponential_euler"


class RigidBody3DSim:
    """GPU-backed FlowableSystem + ConstrainedSystem for rigidbody3d (rigidbody3d/RigidBody3DSim.h:34)."""

    def __init__(self, state, device=0, ctx=None):
        self.ctx = ctx or Context(device)
        self.state = st = state
        lib, h = self.ctx.lib, self.ctx.h
        # a context numbers its meshes in the order they were added over its whole life: the state's own mesh numbers go through the indices add_mesh returns
        self.mesh_index = []
        for mesh in st.meshes:
            idx = C.c_uint32()
            self.ctx.check(lib.sg_rb3d_add_mesh(h, mesh.verts.shape[0], _ptr(mesh.verts), mesh.samples.shape[0], _ptr(mesh.samples), mesh.hull.shape[0], _ptr(mesh.hull),
                                                _ptr(mesh.cell_delta), _ptr(mesh.dims), _ptr(mesh.origin), _ptr(mesh.sdf), C.byref(idx)))
            self.mesh_index.append(int(idx.value))
        geo_mesh = np.ascontiguousarray(st.geo_mesh, dtype=np.uint32).copy()
        for k in range(geo_mesh.shape[0]):
            if int(st.geo_type[k]) == 3 and int(geo_mesh[k]) < len(self.mesh_index):   # SG_GEO_MESH
                geo_mesh[k] = self.mesh_index[int(geo_mesh[k])]
        self.ctx.check(lib.sg_rb3d_set_geometry(h, st.geo_type.shape[0], _ptr(st.geo_type), _ptr(st.geo_r), _ptr(st.geo_half), _ptr(geo_mesh)))
        self.ctx.check(lib.sg_rb3d_set_bodies(h, st.nbodies(), _ptr(st.geo_of_body), _ptr(st.fixed), _ptr(st.m), _ptr(st.I0)))
        self.ctx.check(lib.sg_rb3d_set_gravity(h, _ptr(st.g)))
        self.ctx.check(lib.sg_rb3d_set_planes(h, st.plane_x.shape[0], _ptr(st.plane_x), _ptr(st.plane_n)))
        self.ctx.check(lib.sg_rb3d_set_cylinders(h, st.cyl_r.shape[0], _ptr(st.cyl_x), _ptr(st.cyl_axis), _ptr(st.cyl_r)))
        pp = st.planar_portals
        if pp:
            cat = lambda f: _f64(np.array([f(p) for p in pp], dtype=np.float64))
            mult = np.ascontiguousarray(np.array([p.multiplier for p in pp]), dtype=np.int32)
            self.ctx.check(lib.sg_rb3d_set_portals(h, len(pp), _ptr(cat(lambda p: p.plane_a_x)), _ptr(cat(lambda p: p.plane_a_n)), _ptr(cat(lambda p: p.plane_b_x)),
                                                   _ptr(cat(lambda p: p.plane_b_n)), _ptr(mult)))
        else:
            self.ctx.check(lib.sg_rb3d_set_portals(h, 0, None, None, None, None, None))
        # RigidBody3DState's M: as its constructor formed it until updateMandMinv has run once (SG_MAP_M_UPDATED, include/scisim_b200.h)
        self.m_updated = False

    def name(self):
        return "rigid_body_3d"

    def setMeshSnapshot(self, mesh, record):
        """The bytes RigidBodyTriangleMesh::serialize writes for mesh number `mesh` of this sim's state (rigidbody3d/Geometry/RigidBodyTriangleMesh.cpp:215-232):
        kept by the library and written back verbatim by serializeState."""
        buf = np.frombuffer(record, dtype=np.uint8).copy()
        self.ctx.check(self.ctx.lib.sg_rb3d_set_mesh_snapshot(self.ctx.h, self.mesh_index[int(mesh)], _ptr(buf), buf.shape[0]))

    def updateMandMinv(self, q=None):
        """RigidBody3DState::updateMandMinv (rigidbody3d/RigidBody3DState.cpp:428-462): returns ( I blocks, Iinv blocks ), 9 doubles per
        body, column-major as they sit in the value arrays of M and Minv.  q=None: the device copy of the last flow's / step's q1.
        Every flow after this call multiplies v0 by the updated M, as the reference does from its second step on."""
        n = self.state.nbodies()
        I, Ii = np.zeros(9 * n), np.zeros(9 * n)
        if q is not None:
            q = _f64(q)
            assert q.size == self.nqdofs()
        self.ctx.check(self.ctx.lib.sg_rb3d_update_m_and_minv(self.ctx.h, _ptr(q) if q is not None else None, _ptr(I), _ptr(Ii)))
        self.m_updated = True
        return I, Ii

    # ---- state I/O (rigidbody3d/RigidBody3DState.cpp:586-668) ----
    def serializeState(self, which=1, m_updated=None):
        """bytes of RigidBody3DState::serialize for the device-resident state (spheres and boxes): which = 0 the uploaded ( q0, v0 ), 1 the last flow's
        ( q1, v1 ).  m_updated: the layout of the world-space blocks of M / Minv -- as the constructor stores them (False) or as updateMandMinv leaves them
        (True); None: True for which = 1 (RigidBody3DSim::flow runs updateMandMinv after every map), else what this sim has done so far."""
        if m_updated is None:
            m_updated = True if which == 1 else self.m_updated
        need = C.c_uint64()
        self.ctx.check(self.ctx.lib.sg_rb3d_state_serialize(self.ctx.h, int(which), 1 if m_updated else 0, None, 0, C.byref(need)))
        buf = np.zeros(int(need.value), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sg_rb3d_state_serialize(self.ctx.h, int(which), 1 if m_updated else 0, _ptr(buf), buf.shape[0], C.byref(need)))
        return buf.tobytes()

    @staticmethod
    def deserializeState(blob, ctx, m_updated=True):
        """RigidBody3DState::deserialize: a sim configured from a snapshot, its ( q, v ) uploaded.  m_updated: whether the snapshot was taken from a running
        simulation (its M is updateMandMinv's; every flow of the restored sim then carries SG_MAP_M_UPDATED) or from a freshly constructed state."""
        sim = RigidBody3DSim.__new__(RigidBody3DSim)
        sim.ctx = ctx
        buf = np.frombuffer(blob, dtype=np.uint8).copy()
        ctx.check(ctx.lib.sg_rb3d_state_deserialize(ctx.h, _ptr(buf), buf.shape[0]))
        n = int(np.frombuffer(blob[:4], dtype=np.uint32)[0])
        # the context holds the real tables; the host-side state object only answers nbodies()
        sim.state = RigidBody3DState([1], [1.0], [[0.0, 0.0, 0.0]], [0], [], np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.ones(n), np.ones((n, 3)), [0.0, 0.0, 0.0],
                                     np.zeros((0, 3)), np.zeros((0, 3)))
        sim.m_updated = bool(m_updated)
        sim.mesh_index = []   # the library re-added the snapshot's meshes and keeps their records itself
        return sim

    # ---- portals (rigidbody3d/RigidBody3DSim.cpp:642-663) ----
    def enforcePeriodicBoundaryConditions(self, q):
        """Teleports the centres of mass that left through a portal; returns the new q."""
        q = _f64(q).copy()
        self.ctx.check(self.ctx.lib.sg_rb3d_enforce_portals(self.ctx.h, _ptr(q)))
        return q

    def teleported(self):
        from ._lib import SgTeleported
        t = SgTeleported()
        self.ctx.check(self.ctx.lib.sg_rb3d_teleported(self.ctx.h, C.byref(t)))
        return TeleportedInfo(t, dim=3)

    def nqdofs(self):
        return 12 * self.state.nbodies()

    def nvdofs(self):
        return 6 * self.state.nbodies()

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        q0, v0 = _f64(q0), _f64(v0)
        assert q0.size == self.nqdofs() and v0.size == self.nvdofs()
        q1 = np.empty_like(q0) if q1 is None else q1
        v1 = np.empty_like(v0) if v1 is None else v1
        self.ctx.check(self.ctx.lib.sg_rb3d_flow(self.ctx.h, kind | (SG_MAP_M_UPDATED if self.m_updated else 0), _ptr(q0), _ptr(v0), float(dt), _ptr(q1), _ptr(v1)))
        return q1, v1

    def computeActiveSet(self, q0, qp, v=None, flags=SG_OUT_ALL, copy=True, resident=False):
        """resident=True: (q0, qp) are the input and output of the last flow() on this sim (SG_IN_RESIDENT)."""
        from ._lib import SG_IN_RESIDENT
        q0, qp = _f64(q0), _f64(qp)
        assert q0.size == self.nqdofs() and qp.size == self.nqdofs()
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_rb3d_active_set(self.ctx.h, _ptr(q0), _ptr(qp), int(flags) | (SG_IN_RESIDENT if resident else 0), C.byref(c)))
        return ActiveSet(c, copy=copy)

    def upload(self, q, v):
        q, v = _f64(q), _f64(v)
        self.ctx.check(self.ctx.lib.sg_rb3d_upload(self.ctx.h, _ptr(q), _ptr(v)))

    def meshStats(self):
        """(sample sweeps served from a TMA-staged SDF brick, sweeps that read the grid directly)"""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.ctx.check(self.ctx.lib.sg_rb3d_mesh_stats(self.ctx.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def step(self, umap, dt):
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_rb3d_step(self.ctx.h, umap.kind | (SG_MAP_M_UPDATED if self.m_updated else 0), float(dt), C.byref(c)))
        return int(c.n_candidates), int(c.n_active)

    def fetch(self, flags=SG_OUT_ALL, want_state=True):
        q1 = np.empty(self.nqdofs()) if want_state else None
        v1 = np.empty(self.nvdofs()) if want_state else None
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_rb3d_fetch(self.ctx.h, int(flags), _ptr(q1) if want_state else None, _ptr(v1) if want_state else None, C.byref(c)))
        return q1, v1, ActiveSet(c)


# ---- rigidbody2d -------------------------------------------------------------------------------------
class RigidBody2DState:
    """Static part of rigidbody2d/RigidBody2DState.h: geometry list, per-body geometry index / fixed flag, the 3N mass
    diagonal (m, m, I), gravity, static planes (normals as given)."""

    def __init__(self, geo_type, geo_r, geo_half, geo_of_body, fixed, M, g=(0.0, 0.0), plane_x=None, plane_n=None, planar_portals=None):
        self.planar_portals = list(planar_portals) if planar_portals is not None else []
        self.geo_type = np.ascontiguousarray(geo_type, dtype=np.uint32)
        self.geo_r = _f64(geo_r)
        self.geo_half = _f64(geo_half).reshape(-1, 2)
        self.geo_of_body = np.ascontiguousarray(geo_of_body, dtype=np.uint32)
        self.fixed = np.ascontiguousarray(fixed, dtype=np.uint8)
        self.M = _f64(M)
        self.g = _f64(g)
        self.plane_x = _f64(plane_x if plane_x is not None else np.zeros((0, 2))).reshape(-1, 2)
        self.plane_n = _f64(plane_n if plane_n is not None else np.zeros((0, 2))).reshape(-1, 2)

    def nbodies(self):
        return self.geo_of_body.shape[0]


class RigidBody2DSim:
    """GPU-backed FlowableSystem + ConstrainedSystem for rigidbody2d; use SymplecticEulerMap / VerletMap with it."""

    def __init__(self, state, device=0, ctx=None):
        self.ctx = ctx or Context(device)
        self.state = st = state
        lib, h = self.ctx.lib, self.ctx.h
        self.ctx.check(lib.sg_rb2d_set_geometry(h, st.geo_type.shape[0], _ptr(st.geo_type), _ptr(st.geo_r), _ptr(st.geo_half)))
        self.ctx.check(lib.sg_rb2d_set_bodies(h, st.nbodies(), _ptr(st.geo_of_body), _ptr(st.fixed), _ptr(st.M)))
        self.ctx.check(lib.sg_rb2d_set_gravity(h, _ptr(st.g)))
        self.ctx.check(lib.sg_rb2d_set_planes(h, st.plane_x.shape[0], _ptr(st.plane_x), _ptr(st.plane_n)))
        pp = st.planar_portals
        if pp:
            cat = lambda f: _f64(np.array([f(p) for p in pp], dtype=np.float64))
            arrs = [cat(lambda p: p.plane_a_x), cat(lambda p: p.plane_a_n), cat(lambda p: p.plane_b_x), cat(lambda p: p.plane_b_n),
                    cat(lambda p: p.velocity), cat(lambda p: p.bounds)]
            self.ctx.check(lib.sg_rb2d_set_portals(h, len(pp), *[_ptr(a) for a in arrs]))
        else:
            self.ctx.check(lib.sg_rb2d_set_portals(h, 0, None, None, None, None, None, None))

    def name(self):
        return "rigid_body_2d"

    # ---- portals (rigidbody2d/RigidBody2DSim.cpp:832-874) ----
    def updatePeriodicBoundaryConditionsStartOfStep(self, next_iteration, dt):
        dx = np.zeros(max(1, len(self.state.planar_portals)))
        self.ctx.check(self.ctx.lib.sg_rb2d_update_portals(self.ctx.h, float(next_iteration * dt), _ptr(dx)))
        return dx[: len(self.state.planar_portals)]

    def enforcePeriodicBoundaryConditions(self, q, v):
        q = _f64(q).copy()
        v = _f64(v).copy()
        self.ctx.check(self.ctx.lib.sg_rb2d_enforce_portals(self.ctx.h, _ptr(q), _ptr(v)))
        return q, v

    def teleported(self):
        from ._lib import SgTeleported
        t = SgTeleported()
        self.ctx.check(self.ctx.lib.sg_rb2d_teleported(self.ctx.h, C.byref(t)))
        return TeleportedInfo(t)

    def nqdofs(self):
        return 3 * self.state.nbodies()

    nvdofs = nqdofs

    def _flow(self, kind, q0, v0, dt, q1=None, v1=None):
        q0, v0 = _f64(q0), _f64(v0)
        assert q0.size == self.nqdofs() and v0.size == self.nvdofs()
        q1 = np.empty_like(q0) if q1 is None else q1
        v1 = np.empty_like(v0) if v1 is None else v1
        self.ctx.check(self.ctx.lib.sg_rb2d_flow(self.ctx.h, kind, _ptr(q0), _ptr(v0), float(dt), _ptr(q1), _ptr(v1)))
        return q1, v1

    def computeActiveSet(self, q0, qp, v=None, flags=SG_OUT_ALL, copy=True, resident=False):
        """resident=True: (q0, qp) are the input and output of the last flow() on this sim (SG_IN_RESIDENT)."""
        from ._lib import SG_IN_RESIDENT
        q0, qp = _f64(q0), _f64(qp)
        assert q0.size == self.nqdofs() and qp.size == self.nqdofs()
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_rb2d_active_set(self.ctx.h, _ptr(q0), _ptr(qp), int(flags) | (SG_IN_RESIDENT if resident else 0), C.byref(c)))
        return ActiveSet(c, copy=copy)

    # ---- resident stepping: the state stays on the device between the map and the detection ----
    def upload(self, q, v):
        q, v = _f64(q), _f64(v)
        assert q.size == self.nqdofs() and v.size == self.nvdofs()
        self.ctx.check(self.ctx.lib.sg_rb2d_upload(self.ctx.h, _ptr(q), _ptr(v)))

    def step(self, umap, dt):
        """flow + computeActiveSet on the device copies; returns (number of candidates, number of active contacts)."""
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_rb2d_step(self.ctx.h, umap.kind, float(dt), C.byref(c)))
        return int(c.n_candidates), int(c.n_active)

    def fetch(self, flags=SG_OUT_ALL, want_state=True):
        q1 = np.empty(self.nqdofs()) if want_state else None
        v1 = np.empty(self.nvdofs()) if want_state else None
        c = SgContacts()
        self.ctx.check(self.ctx.lib.sg_rb2d_fetch(self.ctx.h, int(flags), _ptr(q1) if want_state else None, _ptr(v1) if want_state else None, C.byref(c)))
        return q1, v1, ActiveSet(c)

    # ---- state I/O: RigidBody2DState's binary snapshot (rigidbody2d/RigidBody2DState.cpp:485-556) ----
    def serializeState(self, which=1):
        """bytes of RigidBody2DState::serialize for the device-resident state: which = 0 the uploaded ( q0, v0 ), 1 the last flow's ( q1, v1 )."""
        need = C.c_uint64()
        self.ctx.check(self.ctx.lib.sg_rb2d_state_serialize(self.ctx.h, int(which), None, 0, C.byref(need)))
        buf = np.zeros(int(need.value), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.sg_rb2d_state_serialize(self.ctx.h, int(which), _ptr(buf), buf.shape[0], C.byref(need)))
        return buf.tobytes()

    @staticmethod
    def deserializeState(blob, ctx):
        """RigidBody2DState::deserialize: a sim configured from a snapshot, its ( q, v ) uploaded."""
        sim = RigidBody2DSim.__new__(RigidBody2DSim)
        sim.ctx = ctx
        buf = np.frombuffer(blob, dtype=np.uint8).copy()
        ctx.check(ctx.lib.sg_rb2d_state_deserialize(ctx.h, _ptr(buf), buf.shape[0]))
        n, nportals = rb2d_snapshot_counts(blob)
        # the context holds the real tables; the host-side state object only answers nbodies() and how many portals there are
        sim.state = RigidBody2DState([0], [1.0], [[0.0, 0.0]], np.zeros(n, np.uint32), np.zeros(n, np.uint8), np.ones(3 * n), planar_portals=[None] * nportals)
        return sim


def rb2d_snapshot_counts(blob):
    """( bodies, portals ) of a RigidBody2DState snapshot the library has accepted: a walk over the layout of scisim_b200/csrc/sg_rb2d_snapshot.h."""
    i64 = lambda at: int(np.frombuffer(blob[at:at + 8], dtype=np.int64)[0])
    nq = i64(0)
    n = nq // 3
    at = 2 * (8 + 8 * nq)                          # q, v
    at += 2 * (24 + 4 * nq + 4 * (nq + 1) + 8 * nq)  # M, Minv
    at += 8 + n                                    # fixed
    at += 8 + 4 * n                                # geometry indices
    ngeo = i64(at); at += 8
    for _ in range(ngeo):
        t = int(np.frombuffer(blob[at:at + 4], dtype=np.int32)[0])
        at += 4 + (8 if t == 0 else 16)
    nforces = i64(at); at += 8 + nforces * (4 + 16)
    nplanes = i64(at); at += 8 + nplanes * 72
    return n, i64(at)
