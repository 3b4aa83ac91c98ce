"""CPU: after GpuBall2DBackend / GpuRigidBody2DBackend / GpuRigidBody3DBackend::deserializeState the shim's force guard must know the masses and the gravity of
the SNAPSHOT (scisim_b200/host/gpu_backend.cpp, GravityOnlyGuard::configureFromSnapshot): fed with the reference's own *State::serialize bytes (oracle/_ref), it
accepts a system whose force is 0 + m g with the scene's masses and gravity and -- in a child process, since the guard prints and exits as the reference does --
refuses another gravity, other masses, and a stream that ends early.  Host logic only; no GPU."""
import os
import subprocess

import numpy as np
import pytest

from scisim_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "scisim_b200", "host")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_rb2d.so")):
        pytest.skip("oracle/_ref not built (the reference tree is not mounted here)")
    from scisim_b200 import build
    build.build_library()
    subprocess.run(["make", "-C", HOST, "libscisim_b200_host.so"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    exe = str(tmp_path_factory.mktemp("guard") / "guard_harness")
    subprocess.run(["g++", "-O2", "-std=c++14", "-o", exe, os.path.join(ROOT, "tests", "guard_harness.cpp"), "-L" + HOST, "-lscisim_b200_host", "-L" + os.path.join(ROOT, "scisim_b200"), "-lscisim_b200",
                    "-Wl,-rpath," + HOST, "-Wl,-rpath," + os.path.join(ROOT, "scisim_b200")], check=True)
    return exe


def _run(exe, tmp_path, layout, blob, masses, g):
    (tmp_path / "blob").write_bytes(blob)
    (tmp_path / "mg").write_bytes(np.concatenate([np.asarray(masses, np.float64), np.asarray(g, np.float64)]).tobytes())
    return subprocess.run([exe, str(layout), str(tmp_path / "blob"), str(tmp_path / "mg")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def _cases():
    from tests.reference_sim_binding import RefBall2DSim, RefRB2DSim, RefRB3DSim
    s = scenes.ball2d_periodic(300, 181, axes="x")
    s["drum_x"], s["drum_r"] = np.array([[3.0, 4.0]]), np.array([90.0])
    s["g"] = np.array([0.3, -9.81])
    s["m"] = np.random.default_rng(1).uniform(0.5, 2.0, size=300)
    yield 0, RefBall2DSim(s, s["portals"]).serialize_state(), s["m"], s["g"]
    s = scenes.rb2d_periodic(250, 182, axes="xy", lees_edwards=0.4)
    s["g"] = np.array([-0.2, -3.0])
    yield 1, RefRB2DSim(s, s["portals"]).serialize_state(), s["M"][0::3], s["g"]
    s = scenes.rb3d_random_boxes(120, 183, spin=True, nfixed_frac=0.0, nplanes=2)
    yield 2, RefRB3DSim(s).serialize_state(), s["m"], s["g"]


def test_guard_follows_the_snapshot(oracle, harness, tmp_path):
    seen = 0
    for layout, blob, m, g in _cases():
        n = len(m)
        ok = _run(harness, tmp_path, layout, blob, m, g)
        assert ok.returncode == 0 and ok.stdout.strip() == "guard ok n=%d" % n, ok.stdout
        g2 = np.array(g, dtype=np.float64); g2[1] += 1.0e-9
        bad = _run(harness, tmp_path, layout, blob, m, g2)
        assert bad.returncode == 1 and "keep the CPU map" in bad.stdout, bad.stdout
        m2 = np.array(m, dtype=np.float64); m2[n // 2] *= 1.0 + 1.0e-12
        bad = _run(harness, tmp_path, layout, blob, m2, g)
        assert bad.returncode == 1 and "keep the CPU map" in bad.stdout, bad.stdout
        cut = _run(harness, tmp_path, layout, blob[: len(blob) - 9], m, g)
        assert cut.returncode == 1 and "Exiting" in cut.stdout, cut.stdout
        seen += 1
    assert seen == 3


def test_state_snapshot_length_splits_a_sim_stream(oracle, harness, tmp_path):
    """<Sim>::serialize writes the state and then the constraint cache into one stream (Ball2DSim.cpp:810-822 and the twins).  sgh_state_snapshot_length -- the
    same parser the deserializeState wrappers use to put the stream back behind the state -- finds where the reference's own state snapshot ends, whatever
    follows it: here the reference's own ConstraintCache::serialize bytes."""
    import ctypes as C
    from tests.test_constraint_cache_cpu import KINDS, REF, _random_contacts, vp
    host = C.CDLL(os.path.join(HOST, "libscisim_b200_host.so"))
    host.sgh_state_snapshot_length.restype = C.c_uint64
    host.sgh_state_snapshot_length.argtypes = [C.c_int, C.c_void_p, C.c_uint64]
    V = C.c_void_p
    for (layout, blob, m, g), sim in zip(_cases(), ("ball2d", "rb2d", "rb3d")):
        ref = getattr(C.CDLL(os.path.join(REF, "libref_%s.so" % sim)), "ref_%s_cache_roundtrip_ex" % sim)
        ref.restype = C.c_int
        ref.argtypes = [C.c_uint32, V, V, V, C.c_uint32, V, C.c_uint32, V, V, V, V, V, C.c_uint64, V, V, C.c_uint64]
        rng = np.random.default_rng(layout)
        st, si, sj = _random_contacts(sim, rng, 200, 50, 3, True)
        r = rng.normal(size=(200, 2))
        z, none, nb = np.zeros(0, dtype=np.uint32), np.zeros((0, 2)), C.c_uint64(0)
        ref(200, vp(st), vp(si), vp(sj), 2, vp(r), 0, vp(z), vp(z), vp(z), vp(none), None, 0, C.byref(nb), None, 0)
        cache = np.zeros(int(nb.value), dtype=np.uint8)
        ref(200, vp(st), vp(si), vp(sj), 2, vp(r), 0, vp(z), vp(z), vp(z), vp(none), vp(cache), cache.shape[0], C.byref(nb), None, 0)
        assert cache.shape[0] > 200 * 8
        stream = np.frombuffer(blob + cache.tobytes(), dtype=np.uint8).copy()
        assert int(host.sgh_state_snapshot_length(layout, vp(stream), stream.shape[0])) == len(blob)
        assert int(host.sgh_state_snapshot_length(layout, vp(stream), len(blob))) == len(blob)
        assert int(host.sgh_state_snapshot_length(layout, vp(stream), len(blob) - 1)) == 0
        assert int(host.sgh_state_snapshot_length(layout, vp(stream), 5)) == 0
        # the wrappers' stream handling: read the rest, parse, seek back -- what is left in the stream is the cache, whatever came before the state
        for prefix in (0, 11):
            (tmp_path / "stream").write_bytes(b"x" * prefix + stream.tobytes())
            out = subprocess.run([harness, "split", str(layout), str(tmp_path / "stream"), str(prefix)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            assert out.returncode == 0 and out.stdout.strip() == "behind=%d sum=%d consumed=%d" % (cache.shape[0], int(cache.astype(np.uint64).sum()), len(blob)), out.stdout
