"""The C++14 host shim (scisim_b200/host/): builds without a GPU; on a GPU its example program must report exactly what
the Python path computes for the same scene (same C ABI underneath)."""
import math
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "scisim_b200", "host")


def _build():
    from scisim_b200 import build
    build.build_library()
    subprocess.run(["make", "-C", HOST], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def test_host_shim_builds_and_links():
    _build()
    assert os.path.exists(os.path.join(HOST, "libscisim_b200_host.so")) and os.path.exists(os.path.join(HOST, "example_ball2d"))
    syms = subprocess.run(["nm", "-DC", os.path.join(HOST, "libscisim_b200_host.so")], stdout=subprocess.PIPE, text=True).stdout
    for name in ("GpuBall2DBackend::computeActiveSet", "GpuSymplecticEulerMap::flow", "GpuVerletMap::flow", "PairImpulseCache::getCachedConstraint"):
        assert name in syms, name


@pytest.mark.gpu
def test_host_shim_example_matches_python_path(gpu_ctx):
    import scisim_b200 as sb
    _build()
    nx, ny = 50, 37
    out = subprocess.run([os.path.join(HOST, "example_ball2d"), str(nx), str(ny)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    vals = dict(re.findall(r"(\w+)=([-0-9.e+]+)", out))
    n = nx * ny
    b = np.arange(n)
    q0 = np.empty(2 * n)
    q0[0::2] = 0.99 * (b % nx) + 0.001 * np.array([math.sin(12.9898 * k) for k in b])
    q0[1::2] = 0.99 * (b // nx) + 0.001 * np.array([math.cos(78.233 * k) for k in b])
    st = sb.Ball2DState(np.full(n, 0.5), np.ones(n), (0.0, -9.81), [[0.0, -0.5], [-0.5, 0.0]], [[0.0, 1.0], [1.0, 0.0]])
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    q1, v1 = sb.SymplecticEulerMap().flow(q0, np.zeros(2 * n), sim, 1, 1.0e-3)
    a = sim.computeActiveSet(q0, q1)
    assert int(vals["n"]) == n
    assert int(vals["candidates"]) == a.n_candidates
    assert int(vals["ball_ball"]) == a.n_body_body and int(vals["plane"]) == a.n_plane
    assert int(vals["cache_hits"]) == a.n_body_body and float(vals["miss_value"]) == 0.0
    assert float(vals["v1y"]) == v1[1] and float(vals["q1y0"]) == q1[1]
