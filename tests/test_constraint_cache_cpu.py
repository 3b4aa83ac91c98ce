"""CPU: ConstraintCache.  The reference's three ConstraintCache.cpp (ball2d/, rigidbody2d/, rigidbody3d/: std::maps keyed by index pairs,
dispatch by constraint.name(), a miss zeroes r) are compiled unchanged into oracle/_ref and fed lists of constraints built with the reference's own
constraint classes; the host shim's PairImpulseCache (scisim_b200/host/gpu_backend.cpp, bound here through its plain-C calls) and the oracle's
ConstraintCache2D (which the device-side sorted-key join is tested against on the GPU, tests/test_assembly_gpu.py) must return the same impulses for
the same stores and queries: hits, misses, out-of-order stores, a key stored twice (the first impulse stays, as std::map::insert), clear, empty."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "scisim_b200", "host")
REF = os.path.join(ROOT, "oracle", "_ref")

vp = lambda a: a.ctypes.data_as(C.c_void_p)

# sim -> { contact type code of include/scisim_b200.h : ( PairImpulseCache kind, key = ( a, b ) of the reference's map from the contact's ( i, j ) ) }
KINDS = {
    "ball2d": {0: (0, lambda i, j: (i, j)), 2: (1, lambda i, j: (j, i)), 1: (2, lambda i, j: (j, i))},
    "rb3d": {10: (0, lambda i, j: (i, j)), 14: (1, lambda i, j: (j, i)), 17: (2, lambda i, j: (j, i)), 11: (3, lambda i, j: (j, i))},
    "rb2d": {20: (0, lambda i, j: (i, j)), 23: (1, lambda i, j: (j, i)), 22: (2, lambda i, j: (i, j)), 21: (3, lambda i, j: (min(i, j), max(i, j)))},
}
PAIR_TYPES = {"ball2d": (0,), "rb3d": (10, 11), "rb2d": (20, 22, 21)}


def _ref(sim):
    path = os.path.join(REF, "libref_%s.so" % sim)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs the reference tree)")
    lib = C.CDLL(path)
    f = getattr(lib, "ref_%s_cache_roundtrip" % sim)
    f.restype = C.c_int
    f.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return f


def _host():
    from scisim_b200 import build
    build.build_library()
    subprocess.run(["make", "-C", HOST, "libscisim_b200_host.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    lib = C.CDLL(os.path.join(HOST, "libscisim_b200_host.so"))
    lib.sgh_cache_create.restype = C.c_void_p
    lib.sgh_cache_destroy.argtypes = [C.c_void_p]
    lib.sgh_cache_clear.argtypes = [C.c_void_p]
    lib.sgh_cache_empty.argtypes = [C.c_void_p]
    lib.sgh_cache_store.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_uint]
    lib.sgh_cache_lookup.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_uint]
    return lib


def _random_contacts(sim, rng, n, nbodies, nstatic, sort):
    types = np.array(sorted(KINDS[sim].keys()), dtype=np.uint32)
    t = rng.choice(types, size=n)
    i = rng.integers(0, nbodies, size=n).astype(np.uint32)
    j = np.where(np.isin(t, PAIR_TYPES[sim]), rng.integers(0, nbodies, size=n), rng.integers(0, nstatic, size=n)).astype(np.uint32)
    pair = np.isin(t, PAIR_TYPES[sim])
    j = np.where(pair & (j == i), (i + 1) % nbodies, j).astype(np.uint32)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    ordered = pair & (t != 11) & (t != 21)          # the free body comes first, the kinematic one second: those stay as drawn
    i, j = np.where(ordered, lo, i).astype(np.uint32), np.where(ordered, hi, j).astype(np.uint32)
    if sort:
        o = np.lexsort((j, i, t))
        t, i, j = t[o], i[o], j[o]
    return np.ascontiguousarray(t), np.ascontiguousarray(i), np.ascontiguousarray(j)


@pytest.mark.parametrize("sim", ["ball2d", "rb3d", "rb2d"])
@pytest.mark.parametrize("ncomp,sort,unique", [(1, True, True), (2, False, True), (3, False, False)])
def test_host_cache_equals_reference_cache(sim, ncomp, sort, unique):
    ref = _ref(sim)
    host = _host()
    rng = np.random.default_rng({"ball2d": 100, "rb3d": 200, "rb2d": 300}[sim] + ncomp)
    st, si, sj = _random_contacts(sim, rng, 3000, 400, 4, sort)
    if unique:
        key = (st.astype(np.uint64) << np.uint64(50)) | (si.astype(np.uint64) << np.uint64(25)) | sj.astype(np.uint64)
        _, first = np.unique(key, return_index=True)
        first.sort()
        st, si, sj = (np.ascontiguousarray(a[first]) for a in (st, si, sj))
    else:
        assert len({(int(a), int(b), int(c)) for a, b, c in zip(st, si, sj)}) < st.shape[0]   # some keys are stored twice
    ns = st.shape[0]
    r = rng.normal(size=(ns, ncomp))
    # queries: every stored constraint once more (shuffled) + as many random ones (mostly misses)
    qt, qi, qj = _random_contacts(sim, rng, ns, 400, 4, False)
    perm = rng.permutation(ns)
    qt, qi, qj = (np.ascontiguousarray(np.concatenate([a[perm], b])) for a, b in ((st, qt), (si, qi), (sj, qj)))
    nq = qt.shape[0]
    want = np.zeros((nq, ncomp))
    empty = ref(ns, vp(st), vp(si), vp(sj), ncomp, vp(r), nq, vp(qt), vp(qi), vp(qj), vp(want))
    assert empty == 0
    hits = (want != 0.0).any(axis=1)
    assert hits[:ns].all() and 0 < (~hits[ns:]).sum()
    h = host.sgh_cache_create()
    try:
        assert host.sgh_cache_empty(h) == 1
        for k in range(ns):
            kind, keyf = KINDS[sim][int(st[k])]
            a, b = keyf(int(si[k]), int(sj[k]))
            host.sgh_cache_store(h, kind, a, b, vp(np.ascontiguousarray(r[k])), ncomp)
        assert host.sgh_cache_empty(h) == 0
        got = np.full((nq, ncomp), -7.0)
        row = np.zeros(ncomp)
        for k in range(nq):
            kind, keyf = KINDS[sim][int(qt[k])]
            a, b = keyf(int(qi[k]), int(qj[k]))
            host.sgh_cache_lookup(h, kind, a, b, vp(row), ncomp)
            got[k] = row
        assert np.array_equal(got, want)
        host.sgh_cache_clear(h)
        assert host.sgh_cache_empty(h) == 1
        host.sgh_cache_lookup(h, 0, 1, 2, vp(row), ncomp)
        assert np.all(row == 0.0)
    finally:
        host.sgh_cache_destroy(h)
    # the reference's cache with nothing stored: empty, every query a miss
    none = np.zeros((0, ncomp))
    z = np.zeros(0, dtype=np.uint32)
    out = np.full((nq, ncomp), 5.0)
    assert ref(0, vp(z), vp(z), vp(z), ncomp, vp(none), nq, vp(qt), vp(qi), vp(qj), vp(out)) == 1
    assert np.all(out == 0.0)


def test_oracle_cache_equals_reference_cache(oracle):
    """oracle/assembly2d.h ConstraintCache2D (through orc_ball2d_cache_store / lookup on the active sets of two consecutive steps) against
    ball2d/ConstraintCache.cpp fed the same two lists."""
    from scisim_b200 import scenes
    from tests import oracle_binding as ob
    ref = _ref("ball2d")
    s = scenes.ball2d_random(500, 19, nplanes=3, ndrums=1)
    o = ob.Ball2DOracle(s)
    q1, v1 = o.flow(0, s["q"], s["v"], s["dt"])
    a = o.active_set(s["q"], q1, "allpairs")
    na = a["type"].shape[0]
    assert (a["type"] == 0).sum() > 50 and (a["type"] == 2).any() and (a["type"] == 1).any()
    r = np.random.default_rng(3).normal(size=(na, 2))
    o.cache_store(np.ascontiguousarray(r.ravel()), 2)
    q2, v2 = o.flow(0, q1, v1, s["dt"])
    b = o.active_set(q1, q2, "allpairs")
    nb = b["type"].shape[0]
    got, hits = o.cache_lookup(nb, 2)
    u32 = lambda x: np.ascontiguousarray(x, dtype=np.uint32)
    want = np.zeros((nb, 2))
    ref(na, vp(u32(a["type"])), vp(u32(a["i"])), vp(u32(a["j"])), 2, vp(np.ascontiguousarray(r)), nb, vp(u32(b["type"])), vp(u32(b["i"])), vp(u32(b["j"])), vp(want))
    assert np.array_equal(got.reshape(nb, 2), want)
    assert hits == int((want != 0.0).any(axis=1).sum()) and 0 < hits < nb


SIM_CODE = {"ball2d": 0, "rb2d": 1, "rb3d": 2}


@pytest.mark.parametrize("sim", ["ball2d", "rb3d", "rb2d"])
@pytest.mark.parametrize("ncomp,unique", [(1, True), (2, False)])
def test_host_cache_snapshot_equals_reference_cache_snapshot(sim, ncomp, unique):
    """ConstraintCache::serialize / deserialize (ball2d/ConstraintCache.cpp:125-174 and the two rigid-body twins), the second half of <Sim>::serialize: the
    host cache writes the reference's bytes for the same stores (a key stored twice is written once, with its first impulse), reads the reference's bytes back
    and answers like it, and the reference's own cache, filled from the host cache's bytes, answers like the original."""
    path = os.path.join(REF, "libref_%s.so" % sim)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs the reference tree)")
    ref = getattr(C.CDLL(path), "ref_%s_cache_roundtrip_ex" % sim)
    ref.restype = C.c_int
    V = C.c_void_p
    ref.argtypes = [C.c_uint32, V, V, V, C.c_uint32, V, C.c_uint32, V, V, V, V, V, C.c_uint64, V, V, C.c_uint64]
    host = _host()
    host.sgh_cache_serialize.restype = C.c_uint64
    host.sgh_cache_serialize.argtypes = [V, C.c_int, V, C.c_uint64]
    host.sgh_cache_deserialize.argtypes = [V, C.c_int, V, C.c_uint64]
    rng = np.random.default_rng({"ball2d": 500, "rb3d": 600, "rb2d": 700}[sim] + ncomp)
    st, si, sj = _random_contacts(sim, rng, 2000, 300, 4, False)
    if unique:
        key = (st.astype(np.uint64) << np.uint64(50)) | (si.astype(np.uint64) << np.uint64(25)) | sj.astype(np.uint64)
        _, first = np.unique(key, return_index=True)
        first.sort()
        st, si, sj = (np.ascontiguousarray(a[first]) for a in (st, si, sj))
    ns = st.shape[0]
    r = rng.normal(size=(ns, ncomp))
    qt, qi, qj = _random_contacts(sim, rng, ns, 300, 4, False)
    perm = rng.permutation(ns)
    qt, qi, qj = (np.ascontiguousarray(np.concatenate([a[perm], b])) for a, b in ((st, qt), (si, qi), (sj, qj)))
    nq = qt.shape[0]
    # the reference: answers and its own snapshot
    want = np.zeros((nq, ncomp))
    nbytes = C.c_uint64(0)
    ref(ns, vp(st), vp(si), vp(sj), ncomp, vp(r), nq, vp(qt), vp(qi), vp(qj), vp(want), None, 0, C.byref(nbytes), None, 0)
    theirs = np.zeros(int(nbytes.value), dtype=np.uint8)
    ref(ns, vp(st), vp(si), vp(sj), ncomp, vp(r), nq, vp(qt), vp(qi), vp(qj), vp(want), vp(theirs), theirs.shape[0], C.byref(nbytes), None, 0)
    assert (want != 0.0).any(axis=1)[:ns].all()

    def fill(h):
        for k in range(ns):
            kind, keyf = KINDS[sim][int(st[k])]
            a, b = keyf(int(si[k]), int(sj[k]))
            host.sgh_cache_store(h, kind, a, b, vp(np.ascontiguousarray(r[k])), ncomp)

    def answers(h):
        got, row = np.full((nq, ncomp), -7.0), np.zeros(ncomp)
        for k in range(nq):
            kind, keyf = KINDS[sim][int(qt[k])]
            a, b = keyf(int(qi[k]), int(qj[k]))
            host.sgh_cache_lookup(h, kind, a, b, vp(row), ncomp)
            got[k] = row
        return got

    h, h2 = host.sgh_cache_create(), host.sgh_cache_create()
    try:
        fill(h)
        need = int(host.sgh_cache_serialize(h, SIM_CODE[sim], None, 0))
        mine = np.zeros(need, dtype=np.uint8)
        assert int(host.sgh_cache_serialize(h, SIM_CODE[sim], vp(mine), need)) == need
        assert need == theirs.shape[0] and np.array_equal(mine, theirs)
        # the host cache from the reference's bytes
        assert host.sgh_cache_deserialize(h2, SIM_CODE[sim], vp(theirs), theirs.shape[0]) == 1
        assert np.array_equal(answers(h2), want)
        assert int(host.sgh_cache_serialize(h2, SIM_CODE[sim], vp(mine), need)) == need and np.array_equal(mine, theirs)
        # truncated streams are refused and leave an empty cache
        for cut in (3, need // 2, need - 1):
            assert host.sgh_cache_deserialize(h2, SIM_CODE[sim], vp(theirs), cut) == 0
            assert host.sgh_cache_empty(h2) == 1
        # the reference's cache from the host cache's bytes
        again = np.full((nq, ncomp), 3.0)
        z, none = np.zeros(0, dtype=np.uint32), np.zeros((0, ncomp))
        ref(0, vp(z), vp(z), vp(z), ncomp, vp(none), nq, vp(qt), vp(qi), vp(qj), vp(again), None, 0, None, vp(mine), need)
        assert np.array_equal(again, want)
        # an empty cache: one zero count per map
        host.sgh_cache_clear(h)
        assert int(host.sgh_cache_serialize(h, SIM_CODE[sim], None, 0)) == 8 * (3 if sim == "ball2d" else 4)
    finally:
        host.sgh_cache_destroy(h)
        host.sgh_cache_destroy(h2)
