"""The product's ball-ball CCD (scisim_b200/csrc/sg_ccd.h -- the header the kernels include, compiled here for the host) against
the reference's compiled CollisionDetectionUtilities.cpp (oracle/_ref) and against the reference's own expressions on a sweep that
is dense exactly where the verdict flips: grazing contacts, quotients next to 1, zero and tiny coefficients, huge magnitudes."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
vp = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def ccdh(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ccdh") / "libccd_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "ccd_harness.cpp")], check=True)
    lib = C.CDLL(out)
    lib.ccdh_ball_ball.restype = C.c_int
    lib.ccdh_ball_ball.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double]
    lib.ccdh_roots.restype = C.c_int
    lib.ccdh_roots.argtypes = [C.c_double] * 3
    lib.ccdh_roots_verbatim.restype = C.c_int
    lib.ccdh_roots_verbatim.argtypes = [C.c_double] * 3
    lib.ccdh_sweep.restype = C.c_uint64
    lib.ccdh_sweep.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _sweep(lib, c):
    c = np.ascontiguousarray(c, dtype=np.float64)
    first, hits = C.c_uint64(0), C.c_uint64(0)
    bad = lib.ccdh_sweep(c.shape[0], vp(c), C.byref(first), C.byref(hits))
    assert bad == 0, (bad, c[first.value])
    return hits.value


def test_division_free_verdict_equals_the_reference_expressions(ccdh):
    rng = np.random.default_rng(11)
    n = 2_000_000
    # generic coefficients of every sign pattern
    c = np.stack([rng.normal(0, 1, n), rng.normal(0, 1, n), np.abs(rng.normal(0, 1, n)) + 1e-12], axis=1)
    h = _sweep(ccdh, c)
    assert n // 20 < h < n
    # discriminant near zero: c1^2 ~ 4 c2 c0
    c1 = rng.normal(0, 1, n); c2 = np.abs(rng.normal(0, 1, n)) + 1e-6
    c0 = c1 * c1 / (4.0 * c2) * (1.0 + rng.integers(-3, 4, n) * 2.0 ** -52)
    _sweep(ccdh, np.stack([c0, c1, c2], axis=1))
    # approaching pairs whose earlier root sits within a few ulps of 1: 2 c0 ~ -c1 + s
    c1 = -np.abs(rng.normal(0, 1, n)) - 1e-3; c2 = np.abs(rng.normal(0, 1, n)) + 1e-3
    c0 = -(c1 + c2)  # root = 1 exactly in real arithmetic: c2 + c1 + c0 = 0
    for d in range(-4, 5):
        _sweep(ccdh, np.stack([c0 * (1.0 + d * 2.0 ** -52), c1, c2], axis=1))
    # the root itself next to 1 by construction: pick b = -c1 + s, set 2 c0 = nextafter(b, +-) and neighbours
    s = np.abs(rng.normal(0, 1, n)); b = -c1 + s
    for tgt in (b, np.nextafter(b, np.inf), np.nextafter(b, -np.inf), np.nextafter(np.nextafter(b, np.inf), np.inf)):
        c0 = 0.5 * tgt
        c2x = (c1 * c1 - s * s) / (4.0 * c0)
        ok = np.isfinite(c2x) & (c2x > 0)
        _sweep(ccdh, np.stack([c0[ok], c1[ok], c2x[ok]], axis=1))
    # zeros, signed zeros, tiny and huge magnitudes, infinities and NaNs
    vals = np.array([0.0, -0.0, 1.0, -1.0, 2.0 ** -1074, -2.0 ** -1074, 1e-310, -1e-310, 1e-250, 1e-200, 1e-199, 1e-150, 1e-20, 1e20, 1e99, 1e100, 1e101,
                     1e150, 1e200, 1e300, -1e300, np.inf, -np.inf, np.nan, 0.5, 3.0, -3.0, 1e-8, -1e-8], dtype=np.float64)
    c2v = vals[~(vals <= 0.0)]  # c2 is a sum of squares: positive, +inf or NaN
    g = np.array(np.meshgrid(vals, vals, c2v, indexing="ij")).reshape(3, -1).T
    with np.errstate(all="ignore"):
        _sweep(ccdh, g)


def test_ccd_equals_the_compiled_reference(ccdh):
    if not os.path.exists(os.path.join(REFDIR, "libref_ball2d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    ref = C.CDLL(os.path.join(REFDIR, "libref_ball2d.so"))
    ref.ref_ball2d_ccd.restype = C.c_int
    ref.ref_ball2d_ccd.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    cases = []
    for k in range(30000):
        q0a, q0b = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        mode = k % 5
        va = rng.uniform(-3, 3, 2) if mode else np.zeros(2)
        vb = rng.uniform(-3, 3, 2) if mode > 1 else va.copy()
        ra, rb = rng.uniform(0.05, 0.6), rng.uniform(0.05, 0.6)
        if mode == 4:
            # touching at the start or at the end of the step, to the last bit
            d = q0b - q0a
            q0b = q0a + d / np.linalg.norm(d) * (ra + rb)
        cases.append((q0a, q0a + va, ra, q0b, q0b + vb, rb))
    for c in json.load(open(os.path.join(ROOT, "tests", "golden", "ccd_cases.json"))):
        cases.append((np.array(c["q0a"], float), np.array(c["q1a"], float), float(c["ra"]), np.array(c["q0b"], float), np.array(c["q1b"], float), float(c["rb"])))
    hits = 0
    for q0a, q1a, ra, q0b, q1b, rb in cases:
        q0a, q1a, q0b, q1b = (np.ascontiguousarray(x, dtype=np.float64) for x in (q0a, q1a, q0b, q1b))
        hr = ref.ref_ball2d_ccd(vp(q0a), vp(q1a), ra, vp(q0b), vp(q1b), rb, None, None)
        hp = ccdh.ccdh_ball_ball(vp(q0a), vp(q1a), ra, vp(q0b), vp(q1b), rb)
        assert hr == hp, (q0a, q1a, ra, q0b, q1b, rb)
        hits += hr
    assert 1000 < hits < len(cases) - 1000
