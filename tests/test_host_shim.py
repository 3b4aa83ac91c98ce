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
    assert os.path.exists(os.path.join(HOST, "example_rigidbody"))
    assert os.path.exists(os.path.join(HOST, "example_ball2d_multi"))
    for name in ("GpuBall2DMultiBackend::computeActiveSet", "GpuBall2DMultiBackend::flow", "GpuMultiSymplecticEulerMap::flow", "GpuMultiVerletMap::flow", "GravityOnlyGuard::verify",
                 "GpuBall2DBackend::computeActiveSet", "GpuBall2DBackend::setPortals", "GpuBall2DBackend::enforcePeriodicBoundaryConditions", "GpuBall2DBackend::teleportedContacts", "GpuSymplecticEulerMap::flow", "GpuVerletMap::flow", "PairImpulseCache::getCachedConstraint",
                 "GpuRigidBody3DBackend::computeActiveSet", "GpuRigidBody3DBackend::setPortals", "GpuRigidBody3DBackend::teleportedContacts", "GpuRigidBody3DBackend::addMesh", "GpuSplitHamMap::flow", "GpuDMVMap::flow",
                 "GpuRigidBody2DBackend::computeActiveSet", "GpuRigidBody2DBackend::setPortals", "GpuRigidBody2DBackend::teleportedContacts", "GpuRB2DSymplecticEulerMap::flow", "GpuRB2DVerletMap::flow"):
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
    # device-side assembly and impulse cache through the shim: Q of the same active set, every stored impulse found again
    asm = sim.assemble()
    assert int(vals["q_nnz"]) == asm["q_values"].shape[0] and int(vals["dev_cache_ok"]) == 1
    qo, qi, qv = asm["q_outer"].tolist(), asm["q_inner"].tolist(), asm["q_values"].tolist()
    trace = 0.0
    for c in range(asm["n_constraints"]):
        for e in range(qo[c], qo[c + 1]):
            if qi[e] == c:
                trace += qv[e]
    assert float(vals["q_trace"]) == trace


@pytest.mark.gpu
def test_host_shim_rigidbody_example_matches_python_path(gpu_ctx):
    """rigidbody3d (spheres + boxes, DMV) and rigidbody2d (circles + rotated boxes, symplectic Euler) through the C++ shim
    against the Python mirror on identically generated scenes."""
    import scisim_b200 as sb
    _build()
    nx, ny, nz = 7, 5, 4
    out = subprocess.run([os.path.join(HOST, "example_rigidbody"), str(nx), str(ny), str(nz)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    lines = {l.split()[0]: dict(re.findall(r"(\w+)=([-0-9.e+]+)", l)) for l in out.strip().splitlines()}
    # ---- rigidbody3d
    ns = nx * ny * nz
    n = 2 * ns
    b = np.arange(n)
    k = b % ns
    box = b >= ns
    s = np.where(box, 0.95, 0.99)
    x = np.empty((n, 3))
    x[:, 0] = np.where(box, 100.0, 0.0) + s * (k % nx) + 0.001 * np.array([math.sin(12.9898 * i) for i in b])
    x[:, 1] = s * ((k // nx) % ny) + 0.001 * np.array([math.cos(78.233 * i) for i in b])
    x[:, 2] = s * (k // (nx * ny)) + 0.001 * np.array([math.sin(37.719 * i) for i in b])
    q0 = np.concatenate([x.ravel(), np.tile(np.eye(3).ravel(), n)])
    v0 = np.zeros(6 * n)
    st = sb.RigidBody3DState([1, 0], [0.5, 0.0], [[0, 0, 0], [0.5, 0.5, 0.5]], [0, 0], [], box.astype(np.uint32), np.zeros(n, np.uint8), np.ones(n), np.full((n, 3), 0.1),
                             (0.0, -9.81, 0.0), [[0.0, -0.5, 0.0]], [[0.0, 1.0, 0.0]])
    sim = sb.RigidBody3DSim(st, ctx=gpu_ctx)
    q1, v1 = sb.DMVMap().flow(q0, v0, sim, 1, 1.0e-3)
    a = sim.computeActiveSet(q0, q1)
    v = lines["rb3d"]
    assert int(v["n"]) == n and int(v["candidates"]) == a.n_candidates
    assert int(v["sphere_sphere"]) == int((a.type == 10).sum()) > 0 and int(v["body_body"]) == int((a.type == 12).sum()) > 0
    assert int(v["plane_sphere"]) == int((a.type == 14).sum()) > 0 and int(v["plane_box"]) == int((a.type == 15).sum()) > 0
    assert float(v["v1y"]) == v1[1] and float(v["q1y0"]) == q1[1]
    assert abs(float(v["psum"]) - float(a.p.sum())) <= 1e-9 * max(1.0, abs(float(a.p.sum())))
    # ---- rigidbody2d
    nx2, ny2 = 3 * nx, 3 * ny
    ns = nx2 * ny2
    n = 2 * ns
    b = np.arange(n)
    k = b % ns
    box = b >= ns
    q0 = np.empty(3 * n)
    q0[0::3] = np.where(box, 100.0, 0.0) + 0.99 * (k % nx2) + 0.001 * np.array([math.sin(12.9898 * i) for i in b])
    q0[1::3] = 0.99 * (k // nx2) + 0.001 * np.array([math.cos(78.233 * i) for i in b])
    q0[2::3] = np.where(box, 0.1 * k, 0.0)
    M = np.tile([1.0, 1.0, 0.2], n)
    st2 = sb.RigidBody2DState([0, 1], [0.5, 0.0], [[0, 0], [0.5, 0.4]], box.astype(np.uint32), np.zeros(n, np.uint8), M, (0.0, -9.81), [[0.0, -0.5]], [[0.0, 1.0]])
    sim2 = sb.RigidBody2DSim(st2, ctx=gpu_ctx)
    q1, v1 = sim2._flow(0, q0, np.zeros(3 * n), 1.0e-3)
    a = sim2.computeActiveSet(q0, q1)
    v = lines["rb2d"]
    assert int(v["n"]) == n and int(v["candidates"]) == a.n_candidates
    assert int(v["circle_circle"]) == int((a.type == 20).sum()) > 0 and int(v["body_body"]) == int((a.type == 22).sum()) > 0
    assert int(v["plane_circle"]) == int((a.type == 23).sum()) > 0 and int(v["plane_body"]) == int((a.type == 24).sum())
    assert float(v["v1y"]) == v1[1] and float(v["q1y0"]) == q1[1]
    assert abs(float(v["psum"]) - float(a.p.sum())) <= 1e-9 * max(1.0, abs(float(a.p.sum())))


@pytest.mark.gpu
def test_host_shim_multi_gpu_example_matches_single_gpu_python_path(gpu_ctx):
    """GpuBall2DMultiBackend (sg_multi behind the C++ shim): 3 slabs of a randomly numbered scene, on as many GPUs as the box has (one
    GPU: the slabs share it), must report exactly the single-GPU list -- same order -- of the Python path."""
    import scisim_b200 as sb
    import torch
    _build()
    nx, ny, slabs = 60, 45, 3
    ndev = max(1, min(slabs, torch.cuda.device_count()))
    out = subprocess.run([os.path.join(HOST, "example_ball2d_multi"), str(nx), str(ny), str(slabs), str(ndev)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
    vals = dict(re.findall(r"(\w+)=([-0-9.e+]+)", out))
    n = nx * ny
    mult = 7919
    while math.gcd(mult, n) != 1:
        mult += 1
    site = (np.arange(n, dtype=np.int64) * mult) % n
    q0 = np.empty(2 * n)
    q0[0::2] = 0.99 * (site % nx) + 0.001 * np.array([math.sin(12.9898 * k) for k in site])
    q0[1::2] = 0.99 * (site // nx) + 0.001 * np.array([math.cos(78.233 * k) for k in site])
    st = sb.Ball2DState(np.full(n, 0.5), np.ones(n), (0.0, -9.81), [[0.0, -0.5], [-0.5, 0.0]], [[0.0, 1.0], [1.0, 0.0]])
    sim = sb.Ball2DSim(st, ctx=gpu_ctx)
    q1, v1 = sb.SymplecticEulerMap().flow(q0, np.zeros(2 * n), sim, 1, 1.0e-3)
    a = sim.computeActiveSet(q0, q1)
    order_sum = 0
    for t, i, j in zip(a.type.tolist(), a.i.tolist(), a.j.tolist()):
        order_sum = (order_sum * 1000003 + i * 31 + j + t) % 1000000007
    assert int(vals["n"]) == n and int(vals["gpus"]) == slabs and int(vals["partitions"]) >= 1
    assert int(vals["candidates"]) == a.n_candidates and int(vals["ball_ball"]) == a.n_body_body and int(vals["plane"]) == a.n_plane
    assert a.n_body_body > 1000 and int(vals["order_sum"]) == order_sum
    dsum = 0.0
    for d in a.depth.tolist():
        dsum += d
    assert float(vals["depth_sum"]) == dsum
    assert float(vals["v1y"]) == v1[1] and float(vals["q1y0"]) == q1[1]


@pytest.mark.gpu
def test_host_shim_maps_refuse_forces_other_than_the_configured_gravity(gpu_ctx):
    """The GPU maps never call fsys.computeForce; GravityOnlyGuard probes it once and exits (the reference's error convention) when the
    system carries anything but the gravity pushed to the back end."""
    _build()
    ok = subprocess.run([os.path.join(HOST, "example_ball2d"), "20", "10"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert ok.returncode == 0 and "candidates=" in ok.stdout
    bad = subprocess.run([os.path.join(HOST, "example_ball2d"), "20", "10", "3.5"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert bad.returncode != 0 and "keep the CPU map" in bad.stdout and "candidates=" not in bad.stdout
