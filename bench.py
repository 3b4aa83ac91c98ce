#!/usr/bin/env python
"""bench.py -- headline benchmark of the SCISim hot path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" = one pass of the hot path over a fixed state: UnconstrainedMap::flow(q0,v0)->(q1,v1)
followed by ConstrainedSystem::computeActiveSet(q0,q1).  Metric = (candidate pairs + active contacts) per
second, whole job.  Workload at N=1: BASELINE.json configs[1] -- 1000x1000 equal balls on a 0.99-spaced lattice,
gravity, 3 static planes, symplectic Euler.  At N>1 the scene is N such slabs side by side along x (weak
scaling), one slab per rank.

  value     state resident in HBM, CUDA events on the library's stream, L2 flushed between steps
  e2e       the same step through the host-buffer C ABI (sg_ball2d_flow + sg_ball2d_active_set): H2D of q0,v0
            (+ q0,q1 for the active set) from pinned memory and D2H of q1,v1 and the whole active set inside
            the timed region
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement of the reference (oracle/, single thread like the reference's path) on
            one full step of the same scene, timed on this box
"""
import argparse
import json
import os

os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line (NCCL prints its version at INFO/VERSION)
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate+active contact pairs/sec (flow + broad phase + narrow phase, ball2d)"
UNIT = "pairs/s"
NX, NY = 1000, 1000


def scene_for_rank(rank, world):
    """Slab `rank` of a (world*NX) x NY lattice, bodies numbered slab-major; every rank holds the same three planes
    (plane indices are global): floor, left wall of the first slab, right wall of the last slab."""
    from scisim_b200 import scenes
    s = scenes.ball2d_lattice(NX, NY, seed=42 + rank, with_planes=True)
    if world > 1:
        s["q"][0::2] += rank * NX * 0.99
        s["plane_x"][2, 0] += (world - 1) * NX * 0.99
    return s


def workload_name(world):
    if world == 1:
        return "configs[1]: synthetic 2D ball pile, 1M equal-radius balls (1000x1000 lattice, spacing 0.99, r=0.5) under gravity, 3 static planes, symplectic Euler"
    return "configs[1] tiled: %d slabs of 1M balls side by side along x (one slab per GPU)" % world


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML back to back (0.5 ms pause) on a background thread during the timed region
    (same fields as the profiling recipe's nvidia-smi line: clocks.sm, clocks.max.sm, clocks_event_reasons.*)."""

    def __init__(self, device):
        import threading
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.0005)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        self._stop.set()
        self.t.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_name):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_rNN.json, written by profiles/summarize.py); None when no capture covers it."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*.json")))
    if not files:
        return None, None
    data = json.load(open(files[-1]))
    for k, v in data.items():
        if ("sg_" + kernel_name) in k or kernel_name in k:
            return v.get("dram_traffic"), os.path.basename(files[-1])
    return None, os.path.basename(files[-1])


def cpu_reference_step(scene, steps, warmup):
    """The reference's CPU path, 1 thread (its hot path is single-threaded even with USE_OPENMP).  Returns
    (pairs per step, per-step seconds, kind, description).
    kind "reference": broad phase + CCD are the reference's OWN ball2d/SpatialGridDetector.cpp and
    scisim/CollisionDetection/CollisionDetectionUtilities.cpp, compiled unchanged from the reference tree against the Eigen
    stand-in (oracle/_ref/libref_ball2d.so, oracle/Makefile.ref); the map and the swept boxes around them are the oracle's
    restatement.  kind "port": everything is the restatement (oracle/_ref not built)."""
    import ctypes as C
    from tests import oracle_binding as ob
    o = ob.Ball2DOracle(scene)
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_ball2d.so")
    times, pairs = [], 0
    if os.path.exists(ref_path):
        import numpy as np
        ref = C.CDLL(ref_path)
        q0 = np.ascontiguousarray(scene["q"], dtype=np.float64)
        r = np.ascontiguousarray(scene["r"], dtype=np.float64)
        n = r.shape[0]
        # static-geometry contacts (a few thousand, O(N x planes) to find) are counted once with the restatement
        q1, _ = o.flow(0, scene["q"], scene["v"], scene["dt"])
        a = o.active_set(scene["q"], q1, "grid")
        n_static = int((a["type"] != 0).sum())
        for it in range(warmup + steps):
            q1, v1 = o.flow(0, scene["q"], scene["v"], scene["dt"])
            t_flow = float(o.lib.orc_ball2d_seconds_flow(o.h))
            nc, na = C.c_uint64(0), C.c_uint64(0)
            q1c = np.ascontiguousarray(q1)
            t0 = time.perf_counter()
            ref.ref_ball2d_detect(C.c_uint32(n), q0.ctypes.data_as(C.c_void_p), q1c.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), C.byref(nc), C.byref(na), None, C.c_uint64(0))
            t = time.perf_counter() - t0 + t_flow
            pairs = int(nc.value) + int(na.value) + n_static
            if it >= warmup:
                times.append(t)
        assert int(nc.value) == a["candidates"].shape[0] and int(na.value) + n_static == a["type"].shape[0]
        return pairs, times, "reference", ("broad phase + CCD = the reference's own SpatialGridDetector.cpp / CollisionDetectionUtilities.cpp (compiled unchanged against "
                                           "oracle/eigen_standin), map + swept boxes = oracle restatement (the map checked bit for bit against the reference's compiled SymplecticEulerMap.cpp, tests/test_oracle_vs_reference.py)")
    for it in range(warmup + steps):
        q1, v1 = o.flow(0, scene["q"], scene["v"], scene["dt"])
        a = o.active_set(scene["q"], q1, "grid")
        t = a["seconds"] + a["seconds_flow"]
        pairs = a["candidates"].shape[0] + a["type"].shape[0]
        if it >= warmup:
            times.append(t)
    return pairs, times, "port", "reference CPU path restated in oracle/ (its own std::map/std::set data structures)"


def cpu_parallel_step(scene, steps, warmup):
    """NOT reference behaviour (the optional second CPU figure): the same step written for a multi-core CPU
    (oracle/ball2d_parallel.h -- all host threads over bodies, flat grid by counting sort), lists built, results identical to
    the restatement's (tests/test_oracle_parallel.py).  Returns the cpu_baseline_parallel object."""
    from tests import oracle_binding as ob
    o = ob.Ball2DOracle(scene)
    kind = 0 if scene["map"] == "symplectic_euler" else 1
    times, pairs, threads = [], 0, 1
    # one thread per core this process may run on (cgroup / affinity aware), not per core of the machine
    try:
        os.environ["ORC_THREADS"] = str(max(1, min(64, len(os.sched_getaffinity(0)))))
    except AttributeError:
        pass
    for it in range(warmup + steps):
        r = o.parallel_step(kind, scene["q"], scene["v"], scene["dt"], keep_lists=True)
        pairs, threads = r["n_candidates"] + r["n_active"] + r["n_static"], r["threads"]
        if it >= warmup:
            times.append(r["seconds"])
    return {"value": pairs * len(times) / sum(times), "unit": UNIT, "cores": threads, "kind": "port-multithreaded",
            "sample": "%d full steps of the same scene, candidate and contact lists built" % len(times),
            "note": "NOT reference behaviour: SCISim's hot path is single-threaded even with USE_OPENMP (that is cpu_baseline); this is the same step "
                    "written for a multi-core CPU (oracle/ball2d_parallel.h), reported so that the GPU figure is not compared with one core only"}


def cpu_parallel_guarded(steps, warmup):
    """cpu_parallel_step in a child process with a time limit: an optional figure must not be able to take the bench line down."""
    import subprocess
    code = "import json, bench; print(json.dumps(bench.cpu_parallel_step(bench.scene_for_rank(0, 1), %d, %d)))" % (steps, warmup)
    try:
        out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
        if out.returncode != 0:
            return {"unavailable": "child exited with %d: %s" % (out.returncode, out.stderr.strip()[-200:])}
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001 -- reported, never raised
        return {"unavailable": repr(e)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = scene_for_rank(0, 1)
    pairs, times, kind, how = cpu_reference_step(scene, args.steps, min(args.warmup, 1))
    total = sum(times)
    value = pairs * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(1), "bodies": NX * NY, "note": how + "; the reference's hot path is single-threaded even with USE_OPENMP (SURVEY.md F2)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": "full 1M-ball step (flow + spatial-grid broad phase + CCD), %d steps; %s" % (len(times), how)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_per_s": len(times) / total,
        "cpu_baseline_parallel": cpu_parallel_guarded(5, 1),
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run on (and therefore allocate pinned host buffers from) the CPU cores NVML reports as
    local to this GPU, so that N ranks' host<->device copies do not all cross the same socket link.  Returns the number
    of cores bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="N>1 halo exchange: peer-memory mailboxes over NVLink (CUDA IPC) or NCCL all_gather + send/recv")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import numpy as np
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: anything a library prints meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import scisim_b200 as sb
    scene = scene_for_rank(rank, world)
    ctx = sb.Context(local_rank)
    umap = sb.SymplecticEulerMap()
    dt = scene["dt"]
    n = scene["r"].shape[0]

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if world == 1:
        st = sb.Ball2DState(scene["r"], scene["m"], scene["g"], scene["plane_x"], scene["plane_n"], scene["drum_x"], scene["drum_r"])
        sim = sb.Ball2DSim(st, ctx=ctx)
        sim.upload(scene["q"], scene["v"])
        step = lambda: sim.step(umap, dt)
    else:
        # one slab per GPU, ghosts exchanged with the neighbouring slabs every step (scisim_b200/slab.py): by the kernels
        # themselves through peer-mapped mailboxes (default), or by NCCL all_gather + send/recv
        from scisim_b200.slab import Ball2DSlabs, GpuSlabBackend
        backend = GpuSlabBackend(ctx, scene, rank * n, ghost_cap=max(4096, n // 128))
        slabs = Ball2DSlabs(backend, rank, world, dist, transport=args.transport)
        step = lambda: slabs.step(umap.kind, dt)
        args.transport = slabs.transport   # what the ranks agreed on (p2p falls back to nccl when a mailbox cannot be mapped)

    # ---------------- resident path: `value` ----------------
    barrier()                   # also warms the barrier's own collective up before anything is timed
    for _ in range(args.warmup):
        ctx.flush_l2()
        pc, pa = step()
    sampler = ClockSampler(local_rank) if rank == 0 else None   # NVML init happens here, before the ranks line up
    barrier()
    launches0 = ctx.launch_count()
    step_ms = []
    for _ in range(args.steps):
        ctx.flush_l2()          # cold L2 for every timed step (outside the event bracket)
        ctx.timer_begin()
        pc, pa = step()
        step_ms.append(ctx.timer_end())
    barrier()
    gpu_launches = ctx.launch_count() - launches0
    t_local = sum(step_ms) / 1e3
    pairs_local = float(pc + pa)

    # ---------------- per-kernel roofline (same steps, events around every kernel) ----------------
    ctx.profile_enable(True)
    ctx.profile_reset()
    for _ in range(args.steps):
        ctx.flush_l2()
        step()
    prof = ctx.profile()
    ctx.profile_enable(False)

    # ---------------- e2e: host buffers in, host lists out, every step ----------------
    flags = sb.SG_OUT_NORMALS | sb.SG_OUT_POINTS | sb.SG_OUT_DEPTHS
    e2e_steps = max(3, min(args.steps, 10))
    q0h, v0h = ctx.pinned((2 * n,)), ctx.pinned((2 * n,))
    q1h, v1h = ctx.pinned((2 * n,)), ctx.pinned((2 * n,))
    q0h[:] = scene["q"]; v0h[:] = scene["v"]
    if world == 1:
        for it in range(2 + e2e_steps):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            sim._flow(umap.kind, q0h, v0h, dt, q1h, v1h)
            a = sim.computeActiveSet(q0h, q1h, flags=flags, copy=False, resident=True)   # (q0, q1) of the flow just done, as ImpactMap::flow passes them
        ctx.synchronize()
        t_e2e_local = time.perf_counter() - t0
        h2d = 2 * 2 * n * 8
        n_act, e2e_api = a.n_active, "sg_ball2d_flow (q0,v0 up; q1,v1 down) + sg_ball2d_active_set(SG_IN_RESIDENT: the flow's q0,q1 stay on the device; contacts down), pinned host buffers, wall clock"
        assert a.n_active == pa and a.n_candidates == pc
    else:
        import ctypes as C
        from scisim_b200._lib import SgContacts
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        for it in range(2 + e2e_steps):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            ctx.check(ctx.lib.sg_ball2d_upload(ctx.h, vp(q0h), vp(v0h)))
            slabs.step(umap.kind, dt)
            c = SgContacts()
            ctx.check(ctx.lib.sg_ball2d_fetch(ctx.h, flags, vp(q1h), vp(v1h), C.byref(c)))
        ctx.synchronize()
        t_e2e_local = time.perf_counter() - t0
        h2d = 2 * 2 * n * 8
        n_act, e2e_api = int(c.n_active), "sg_ball2d_upload + slab step (%s halo) + sg_ball2d_fetch, pinned host buffers, wall clock" % args.transport
    d2h = 2 * 2 * n * 8 + n_act * (4 + 4 + 4 + 16 + 16 + 8)
    # the sampler ran from just before the timed steps to here (timed region, per-kernel pass, e2e pass: all under load)
    clocks = sampler.stop() if sampler else None

    # ---------------- reduce over ranks: max time, summed work ----------------
    if world > 1:
        tt = torch.tensor([t_local, t_e2e_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ww = torch.tensor([pairs_local, float(pc), float(pa)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        t_max, t_e2e = float(tt[0]), float(tt[1])
        pairs_all = float(ww[0])
    else:
        t_max, t_e2e, pairs_all = t_local, t_e2e_local, pairs_local

    if rank == 0:
        peak, peak_src = measured_peak()
        total_ms = sum(v[1] for v in prof.values())
        top = max(prof.items(), key=lambda kv: kv[1][1])
        name, (nl, ms, by) = top
        achieved = (by / nl) / (ms / nl * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(name)
        kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps, "share": v[1] / total_ms,
                       "alg_GBps": (v[2] / (v[1] * 1e-3) / 1e9) if v[1] > 0 else None} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        line = {
            "metric": METRIC, "value": pairs_all * args.steps / t_max, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world), "bodies_per_gpu": n, "candidates_per_gpu": pc, "active_per_gpu": pa,
                       "l2": "384 MB buffer overwritten before every timed step (L2 flush)", "timing": "CUDA events on the library stream, per step, summed; max over ranks",
                       "parallelism": ("%d x-slabs, 1 process per GPU, ghost bodies exchanged with +-1 neighbours each step (%s), pair owned by the rank of its lower index" % (world, "peer-memory mailboxes over NVLink, no collective in the step" if args.transport == "p2p" else "NCCL all_gather + send/recv")) if world > 1 else "single GPU"},
            "host_cores_bound": numa,
            "steps_per_s": args.steps / t_max,
            "step_ms_rank0": [round(x, 4) for x in step_ms],
            "clocks": clocks,
            "gpu_launches": gpu_launches,
            "e2e": {"value": pairs_all * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / e2e_steps,
                    "api": e2e_api},
            "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": by / nl,
                         "peak_source": peak_src, "share_of_step": ms / total_ms, "timed": "separate pass of the same %d steps with CUDA events around every kernel" % args.steps,
                         "kernels": kernels},
        }
        if not args.no_cpu_baseline:
            pairs_cpu, times, kind, how = cpu_reference_step(scene_for_rank(0, 1), 3, 1)
            v = pairs_cpu * len(times) / sum(times)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "3 full steps of the same 1M-ball scene; %s; the reference's hot path is single-threaded" % how}
            line["cpu_baseline_parallel"] = cpu_parallel_guarded(5, 1)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
