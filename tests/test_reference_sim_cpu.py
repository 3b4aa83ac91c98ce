"""CPU: the reference's OWN simulation classes, compiled unchanged into oracle/_ref (Ball2DSim.cpp + Ball2DState.cpp, RigidBody3DSim.cpp + RigidBody3DState.cpp
and everything they call; oracle/ref_shims/ref_*_sim.cpp), run here and compared with the oracle:

  * <Sim>::computeActiveSet( q0, q1, v ) as a whole -- broad phase, narrow phases, the portal branch with its teleported collisions, static geometry, in the
    order the reference emits the constraints -- against the oracle's active set: types, indices, normals, points, depths, element by element, bit for bit;
  * Ball2DSim::flow( call_back, iteration, dt, umap ) over many steps (portal motion and periodic boundaries included) against the oracle's
    flow / update_portals / enforce_portals sequence.

This pins the oracle's GLUE (emission order, dispatch, kinematic rules, the teleported-collision set), which the per-class pins of
tests/test_oracle_vs_reference.py leave restated.  The GPU parity tests compare the product with this same oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from scisim_b200 import scenes
from tests import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)


def _lib(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    return C.CDLL(path)


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


def _same_active_set(got, want, dim):
    assert got["type"].shape[0] == want["type"].shape[0], (got["type"].shape[0], want["type"].shape[0])
    for k in ("type", "i", "j"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["n"], want["n"]) and np.array_equal(got["p"], want["p"])
    assert np.array_equal(got["depth"], want["depth"], equal_nan=True)


@pytest.mark.parametrize("seed,n,nplanes,ndrums,kind", [(61, 1500, 3, 1, 0), (62, 900, 2, 2, 1), (63, 2, 1, 0, 0), (64, 2500, 4, 0, 1)])
def test_ball2d_sim_compute_active_set(oracle, seed, n, nplanes, ndrums, kind):
    """Ball2DSim::computeActiveSet (Ball2DSim.cpp:151-173, 553-608, 730-762) on messy scenes: ball-ball (swept boxes + CCD), then drums, then planes."""
    s = scenes.ball2d_random(n, seed, nplanes=nplanes, ndrums=ndrums)
    o = ob.Ball2DOracle(s)
    ref = RefBall2DSim(s)
    q, v = s["q"].copy(), s["v"].copy()
    seen = set()
    for step in range(3):
        q1, v1 = o.flow(kind, q, v, s["dt"])
        want = o.active_set(q, q1, "grid")
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        seen |= set(int(t) for t in got["type"])
        q, v = q1, v1
    assert 0 in seen or n < 10
    assert (2 in seen or nplanes == 0) and (1 in seen or ndrums == 0)


@pytest.mark.parametrize("name,steps", [("pool_break_ten_deep", 40), ("different_friction", 12)])
def test_ball2d_sim_on_bundled_scenes(oracle, name, steps):
    """BASELINE configs[0]: the reference's own scenes through the reference's own Ball2DSim -- flow( SymplecticEulerMap ) then computeActiveSet, step after
    step -- against the oracle.  (Unconstrained: no contact response on this path, as in the GPU step.)"""
    s = scenes.ball2d_asset(name)
    den = 10080 if name == "different_friction" else 10
    assert s["dt"] == 1.0 / den
    o = ob.Ball2DOracle(s)
    ref = RefBall2DSim(s)
    q, v = s["q"].copy(), s["v"].copy()
    total = 0
    for it in range(1, steps + 1):
        q1, v1 = o.flow(0, q, v, s["dt"])
        want = o.active_set(q, q1, "grid")
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        total += got["type"].shape[0]
        rq, rv = ref.flow(0, it, 1, den)
        assert np.array_equal(rq, q1) and np.array_equal(rv, v1)
        q, v = q1, v1
    assert total > 0 or name == "different_friction"   # (6 079 balls falling side by side: no contacts in the first steps, see tests/test_config1_cpu.py)


@pytest.mark.parametrize("axes,le,oblique,seed", [("x", 0.0, False, 71), ("xy", 0.0, False, 72), ("xy", 0.8, False, 73), ("y", -1.3, True, 74), ("xy", 0.5, True, 75)])
def test_ball2d_sim_with_portals(oracle, axes, le, oblique, seed):
    """The portal branch (Ball2DSim.cpp:159-166, 327-546, 610-728): teleported boxes, the TeleportedCollision set, BallBallConstraint /
    KinematicKickBallBallConstraint from teleported centres; and Ball2DSim::flow with moving (Lees-Edwards) portals and the periodic wrap, 8 steps."""
    s = scenes.ball2d_periodic(700, seed, axes=axes, lees_edwards=le, oblique=oblique)
    o = ob.Ball2DOracle(s)
    o.set_portals(s["portals"])
    ref = RefBall2DSim(s, s["portals"])
    q, v = s["q"].copy(), s["v"].copy()
    q, v = o.enforce_portals(q, v)
    ref.set_state(q, v)
    den = 100
    assert s["dt"] == 1.0 / den
    seen = set()
    for it in range(1, 9):
        o.update_portals(it * s["dt"])
        q1, v1 = o.flow(1, q, v, s["dt"])
        want = o.active_set_portals(q, q1, "grid")
        assert want is not None
        # the reference's computeActiveSet reads the portals' current offsets: advance its own copy through its flow first
        rq, rv = ref.flow(1, it, 1, den)
        q1w, v1w = o.enforce_portals(q1, v1)
        assert np.array_equal(rq, q1w) and np.array_equal(rv, v1w)
        got = ref.active_set(q, q1)
        _same_active_set(got, want, 2)
        seen |= set(int(t) for t in got["type"])
        q, v = q1w, v1w
    assert 0 in seen and (3 in seen or 4 in seen) and (le == 0.0 or 4 in seen) and (axes != "xy" or 3 in seen)
