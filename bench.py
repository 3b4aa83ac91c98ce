#!/usr/bin/env python
"""bench.py -- headline benchmark of the SCISim hot path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 3|2|1]

A "step" = one pass of the hot path over a fixed state: UnconstrainedMap::flow(q0,v0)->(q1,v1)
followed by ConstrainedSystem::computeActiveSet(q0,q1).  Metric = (candidate pairs + active contacts) per
second, whole job.

Workload (default, every N): BASELINE.json configs[2] -- the north-star scene: 16 777 216 polydisperse balls (radii
log-uniform in [0.25, 1], area fraction 0.55, random numbering), 4 walls, Verlet.  It fits one B200, so N = 1 runs it
whole; N > 1 cuts it into N equal-count x-quantile slabs, one per rank ("scaling": "strong"), ghosts exchanged with
the neighbouring slabs every step through peer-mapped mailboxes (NVLink), pairs owned by the rank of their lower
global index.  --config 2 (N = 1 only) is configs[1], the 1M-ball pile round 1 was quoted on; --config 1 is the
reference's own bundled scene (tests/golden/ball2d_assets.npz).

  value     state resident in HBM, CUDA events on the library's stream, L2 flushed between steps
  e2e       the same step through the host-buffer C ABI: H2D of q0,v0 from pinned memory and D2H of q1,v1 and the
            whole active set inside the timed region
  roofline  dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the reference's own CPU path (its Ball2DSim::flow + Ball2DSim::computeActiveSet, compiled unchanged, single
            thread like the reference) on a bounded sample of the same workload, timed on this box
  parity_check  a reduced-size scene of the same kind through the same code path at this N (partition, halo exchange,
            merge included), compared with the CPU oracle bit for bit
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate+active contact pairs/sec (flow + broad phase + narrow phase, ball2d)"
UNIT = "pairs/s"
N_CONFIG3 = 1 << 24
N_CPU_SAMPLE = 1 << 19      # bodies of the CPU arms' bounded sample of config 3 (the reference needs ~6 us per body per step on this scene)
N_PARITY = 200_000


def make_scene(config, n=None):
    from scisim_b200 import scenes
    if config == 3:
        return scenes.ball2d_gas(n=n or N_CONFIG3)
    if config == 2:
        return scenes.ball2d_lattice(1000, 1000, seed=42, with_planes=True)
    if config == 1:
        return scenes.ball2d_asset("different_friction")
    raise SystemExit("unknown --config %r" % config)


def workload_name(config):
    return {3: "configs[2]: synthetic polydisperse 2D ball gas, 16 777 216 balls (radius ratio 1:4, log-uniform radii, area fraction 0.55, random numbering), 4 walls, Verlet",
            2: "configs[1]: synthetic 2D ball pile, 1M equal-radius balls (1000x1000 lattice, spacing 0.99, r=0.5) under gravity, 3 static planes, symplectic Euler",
            1: "configs[0]: bundled assets/ball2d scene tests_python_serialization/different_friction.xml (6 079 balls, 3 static planes; portal ignored), integrator forced to symplectic_euler, dt = 1/10080"}[config]


def config_dict(config, scene_n):
    """Identical in both arms (the driver compares it)."""
    return {"workload": workload_name(config), "bodies": int(scene_n)}


def map_kind(scene):
    return 0 if scene["map"] == "symplectic_euler" else 1


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


def ncu_traffic(kernel_name, config, n_local):
    """DRAM bytes (read + write) per launch of the dominant kernel, from the committed `ncu --set full` capture of this workload
    (profiles/ncu_r2_c3_2m.json etc., written by profiles/summarize_r2.py): the capture's dram__bytes_read.sum + dram__bytes_write.sum of
    that kernel, scaled by bodies per launch here / bodies in the capture (ncu cannot replay the 16M-ball scene: it saves and restores
    the 11 GB of device memory around every pass).  None when no capture covers the workload."""
    name = {3: "ncu_r2_c3_2m.json", 2: "ncu_r2_c2.json"}.get(config)
    p = os.path.join(ROOT, "profiles", name) if name else None
    if p is None or not os.path.exists(p):
        return None, None
    data = json.load(open(p))
    hit = None
    for k, v in data.get("ncu_kernels", {}).items():
        if ("sg_" + kernel_name) in k and "dram_read_MB" in v:
            hit = v   # the last matching kernel of the warm step (config 2 holds both pass-1 kernels: the staged one comes later)
    if hit is None:
        return None, name
    scale = float(n_local) / float(data["bodies"])
    return (hit["dram_read_MB"] + hit.get("dram_write_MB", 0.0)) * 1e6 * scale, "%s (capture of %d bodies, scaled x%.3g)" % (name, data["bodies"], scale)


def cpu_reference_step(scene, steps, warmup):
    """The reference's CPU path, 1 thread (its hot path is single-threaded even with USE_OPENMP).  Returns
    (pairs per step, per-step seconds, kind, description).
    kind "reference": the step IS the reference's own code -- Ball2DSim::flow( call_back, iteration, dt, umap ) followed by
    Ball2DSim::computeActiveSet( q0, q1, v ) of ball2d/Ball2DSim.cpp, with Ball2DState.cpp, the maps, SpatialGridDetector.cpp,
    CollisionDetectionUtilities.cpp and the constraint classes, all compiled unchanged from the reference tree against the Eigen
    stand-in (oracle/_ref/libref_ball2d.so, oracle/Makefile.ref, oracle/ref_shims/ref_ball2d_sim.cpp): its std::map grid, its std::set
    of pairs, one heap-allocated Constraint per contact, the drum and plane loops.  Timed inside the library around those two calls.
    The candidate count (which the reference's API does not return) comes from one untimed call of its detector.
    kind "port": everything is the restatement (oracle/_ref not built)."""
    import ctypes as C
    kind = map_kind(scene)
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_ball2d.so")
    times, pairs = [], 0
    if os.path.exists(ref_path):
        import numpy as np
        from fractions import Fraction
        from tests.reference_sim_binding import RefBall2DSim
        ref = C.CDLL(ref_path)
        sim = RefBall2DSim(scene)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        q0 = np.ascontiguousarray(scene["q"], dtype=np.float64)
        v0 = np.ascontiguousarray(scene["v"], dtype=np.float64)
        r = np.ascontiguousarray(scene["r"], dtype=np.float64)
        n = r.shape[0]
        dt = Fraction(scene["dt"]).limit_denominator(1 << 40)
        assert float(dt.numerator) / float(dt.denominator) == scene["dt"]   # scalar( Rational ) = numerator / denominator in double
        ref.ref_ball2d_sim_step_timed.restype = C.c_double
        ref.ref_ball2d_sim_step_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]
        nc = None
        for it in range(warmup + steps):
            na, tf = C.c_uint64(0), C.c_double(0.0)
            t = float(ref.ref_ball2d_sim_step_timed(sim.h, vp(q0), vp(v0), kind, 1, dt.numerator, dt.denominator, C.byref(na), C.byref(tf)))
            if nc is None:
                sim.set_state(q0, v0)
                q1, _ = sim.flow(kind, 1, dt.numerator, dt.denominator)
                c, a = C.c_uint64(0), C.c_uint64(0)
                ref.ref_ball2d_detect(C.c_uint32(n), vp(q0), vp(np.ascontiguousarray(q1)), vp(r), C.byref(c), C.byref(a), None, C.c_uint64(0))
                nc = int(c.value)
            pairs = nc + int(na.value)
            if it >= warmup:
                times.append(t)
        return pairs, times, "reference", ("the reference's own Ball2DSim::flow + Ball2DSim::computeActiveSet (ball2d/Ball2DSim.cpp, Ball2DState.cpp, the maps, SpatialGridDetector.cpp, "
                                           "CollisionDetectionUtilities.cpp and the constraint classes compiled unchanged against oracle/eigen_standin; -O3 -DNDEBUG, timed around the two calls "
                                           "inside the library; static-geometry contacts timed and counted, as in the GPU arm)")
    from tests import oracle_binding as ob
    o = ob.Ball2DOracle(scene)
    for it in range(warmup + steps):
        q1, v1 = o.flow(kind, scene["q"], scene["v"], scene["dt"])
        a = o.active_set(scene["q"], q1, "grid")
        t = a["seconds"] + a["seconds_flow"]
        pairs = a["candidates"].shape[0] + a["type"].shape[0]
        if it >= warmup:
            times.append(t)
    return pairs, times, "port", "reference CPU path restated in oracle/ (its own std::map/std::set data structures)"


def cpu_sample_scene(config):
    """The bounded sample the CPU arms run: configs 1 and 2 whole, config 3 at N_CPU_SAMPLE bodies (same generator, same
    density and radius distribution)."""
    if config == 3:
        return make_scene(3, N_CPU_SAMPLE), "config 3 at %d balls (same generator: same density, radius distribution, walls and map; the 16M scene's std::set alone would need > 4 GB and minutes per step)" % N_CPU_SAMPLE
    return make_scene(config), "the whole scene"


def cpu_parallel_step(scene, steps, warmup):
    """NOT reference behaviour (the optional second CPU figure): the same step written for a multi-core CPU
    (oracle/ball2d_parallel.h -- all host threads over bodies, flat grid by counting sort), lists built, results identical to
    the restatement's (tests/test_oracle_parallel.py).  Returns the cpu_baseline_parallel object."""
    from tests import oracle_binding as ob
    o = ob.Ball2DOracle(scene)
    kind = map_kind(scene)
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
            "sample": "%d full steps of the CPU sample scene, candidate and contact lists built" % len(times),
            "note": "NOT reference behaviour: SCISim's hot path is single-threaded even with USE_OPENMP (that is cpu_baseline); this is the same step "
                    "written for a multi-core CPU (oracle/ball2d_parallel.h), reported so that the GPU figure is not compared with one core only"}


def cpu_parallel_guarded(config, steps, warmup):
    """cpu_parallel_step in a child process with a time limit: an optional figure must not be able to take the bench line down."""
    import subprocess
    code = "import json, bench; print(json.dumps(bench.cpu_parallel_step(bench.cpu_sample_scene(%d)[0], %d, %d)))" % (config, steps, warmup)
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
    scene, sample = cpu_sample_scene(args.config)
    full_n = N_CONFIG3 if args.config == 3 else scene["r"].shape[0]
    warm = min(args.warmup, 1)
    pairs, times, kind, how = cpu_reference_step(scene, args.steps, warm)
    total = sum(times)
    value = pairs * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.config, full_n),
        "note": how + "; the reference's hot path is single-threaded even with USE_OPENMP (SURVEY.md F2); warm-up steps actually run: %d" % warm,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": "%s, %d bodies, %d timed steps (flow + computeActiveSet); %s" % (sample, scene["r"].shape[0], len(times), how)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_per_s": len(times) / total,
        "cpu_baseline_parallel": cpu_parallel_guarded(args.config, 3, 1),
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


def parity_check(config, world, rank, dist, ctx, transport):
    """A reduced-size scene of the same kind through the SAME code path the timed steps take at this N -- on N > 1:
    x-quantile partition, halo exchange, per-rank detection, merge into the reference's order -- against the CPU oracle."""
    import numpy as np
    import scisim_b200 as sb
    s = make_scene(3, N_PARITY) if config == 3 else make_scene(config)
    kind = map_kind(s)
    umap = sb.SymplecticEulerMap() if kind == 0 else sb.VerletMap()
    if world == 1:
        st = sb.Ball2DState(s["r"], s["m"], s["g"], s["plane_x"], s["plane_n"], s["drum_x"], s["drum_r"])
        sim = sb.Ball2DSim(st, ctx=ctx)
        sim.upload(s["q"], s["v"])
        sim.step(umap, s["dt"])
        q1, v1, a = sim.fetch()
        got = {"q1": q1, "v1": v1, "candidates": a.candidates, "type": a.type, "i": a.i, "j": a.j, "n": a.n, "p": a.p, "depth": a.depth}
        parts = 1
    else:
        from scisim_b200.slab import Ball2DSlabSim, GpuSlabBackend
        factory = lambda sl, gids, lim, cap: GpuSlabBackend(ctx, sl, 0, cap, gids=gids, x_limits=lim)
        sim = Ball2DSlabSim(s, rank, world, dist, factory, transport=transport)
        sim.upload(s["q"], s["v"])
        sim.step(kind, s["dt"])
        got = sim.gather_merged(0)
        parts = sim.n_partitions
        sim.backend.disconnect()
    if rank != 0:
        return None
    from tests import oracle_binding as ob
    o = ob.Ball2DOracle(s)
    rq1, rv1 = o.flow(kind, s["q"], s["v"], s["dt"])
    ref = o.active_set(s["q"], rq1, "grid")
    bad = [k for k in ("q1", "v1") if not np.array_equal(got[k], (rq1, rv1)[k == "v1"])]
    bad += [k for k in ("candidates", "type", "i", "j", "n", "p") if not np.array_equal(got[k], ref[k])]
    if not np.array_equal(got["depth"], ref["depth"], equal_nan=True):
        bad.append("depth")
    return {"ok": not bad, "mismatch": bad, "bodies": int(s["r"].shape[0]), "candidates": int(ref["candidates"].shape[0]), "active": int(ref["type"].shape[0]), "partitions": parts,
            "what": "q1, v1, candidate list and active set (type, i, j, n, p, depth) of a reduced-size scene of this workload, through the same path as the timed steps at this N, bit-identical to the CPU oracle in the reference's order"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3], help="3: BASELINE configs[2] (16M polydisperse balls; default, any N); 2: configs[1] (1M pile, N = 1); 1: configs[0] (bundled scene, N = 1)")
    ap.add_argument("--bodies", type=int, default=0, help="config 3 only: override the body count (for experiments; the headline is 16 777 216)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config-2 side measurement at N = 1")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="N>1 halo exchange: peer-memory mailboxes over NVLink (CUDA IPC) or NCCL all_gather + send/recv")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.config != 3:
        raise SystemExit("--config %d is a single-GPU workload" % args.config)
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
    scene = make_scene(args.config, args.bodies or None)
    n_total = scene["r"].shape[0]
    ctx = sb.Context(local_rank)
    kind = map_kind(scene)
    umap = sb.SymplecticEulerMap() if kind == 0 else sb.VerletMap()
    dt = scene["dt"]

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    parity = None
    if not args.no_parity:
        parity = parity_check(args.config, world, rank, dist, ctx, args.transport)
        barrier()

    if world == 1:
        st = sb.Ball2DState(scene["r"], scene["m"], scene["g"], scene["plane_x"], scene["plane_n"], scene["drum_x"], scene["drum_r"])
        sim = sb.Ball2DSim(st, ctx=ctx)
        sim.upload(scene["q"], scene["v"])
        step = lambda: sim.step(umap, dt)
        n_local = n_total
        halo = (0, 0)
    else:
        # one x-quantile slab per GPU, ghosts exchanged with the neighbouring slabs every step (scisim_b200/slab.py): by the
        # kernels themselves through peer-mapped mailboxes (default), or by NCCL all_gather + send/recv
        from scisim_b200.slab import Ball2DSlabSim, GpuSlabBackend
        factory = lambda sl, gids, lim, cap: GpuSlabBackend(ctx, sl, 0, cap, gids=gids, x_limits=lim)
        slabs = Ball2DSlabSim(scene, rank, world, dist, factory, transport=args.transport)
        slabs.upload(scene["q"], scene["v"])
        step = lambda: slabs.step(kind, dt)
        args.transport = slabs.transport   # what the ranks agreed on (p2p falls back to nccl when a mailbox cannot be mapped)
        n_local = int(slabs.gids.shape[0])

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
    if world > 1:
        halo = slabs.driver.last_halo

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
    e2e_steps = max(3, min(args.steps, 5))
    q0h, v0h = ctx.pinned((2 * n_local,)), ctx.pinned((2 * n_local,))
    q1h, v1h = ctx.pinned((2 * n_local,)), ctx.pinned((2 * n_local,))
    if world == 1:
        q0h[:] = scene["q"]; v0h[:] = scene["v"]
        for it in range(1 + e2e_steps):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            sim._flow(umap.kind, q0h, v0h, dt, q1h, v1h)
            a = sim.computeActiveSet(q0h, q1h, flags=flags, copy=False, resident=True)   # (q0, q1) of the flow just done, as ImpactMap::flow passes them
        ctx.synchronize()
        t_e2e_local = time.perf_counter() - t0
        n_act, e2e_api = a.n_active, "sg_ball2d_flow (q0,v0 up; q1,v1 down) + sg_ball2d_active_set(SG_IN_RESIDENT: the flow's q0,q1 stay on the device; contacts down), pinned host buffers, wall clock"
        assert a.n_active == pa and a.n_candidates == pc
        # the same call as SCISim's shim makes it: the reference's constraint constructors take (type, i, j) and the host
        # state, so normals / points / depths need not cross the bus -- reported beside the headline e2e, which downloads everything
        for it in range(1 + e2e_steps):
            if it == 1:
                ctx.synchronize()
                t0 = time.perf_counter()
            sim._flow(umap.kind, q0h, v0h, dt, q1h, v1h)
            a2 = sim.computeActiveSet(q0h, q1h, flags=0, copy=False, resident=True)
        ctx.synchronize()
        t_shim = time.perf_counter() - t0
        assert a2.n_active == pa
        e2e_shim = {"value": pairs_local * e2e_steps / t_shim, "unit": UNIT, "ms_per_step": 1e3 * t_shim / e2e_steps, "h2d_bytes_per_step": 2 * 2 * n_local * 8,
                    "d2h_bytes_per_step": 2 * 2 * n_local * 8 + int(a2.n_active) * 12,
                    "api": "as e2e with out_flags = 0: (type, i, j) of every contact come back, no normals / points / depths -- what the SCISim-side shim needs to call the reference's constraint constructors"}
    else:
        e2e_shim = None
        import ctypes as C
        from scisim_b200._lib import SgContacts
        vp = lambda x: x.ctypes.data_as(C.c_void_p)
        q0h[:] = scene["q"].reshape(-1, 2)[slabs.gids].ravel(); v0h[:] = scene["v"].reshape(-1, 2)[slabs.gids].ravel()
        for it in range(1 + e2e_steps):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            ctx.check(ctx.lib.sg_ball2d_upload(ctx.h, vp(q0h), vp(v0h)))
            slabs.step(kind, dt)
            c = SgContacts()
            ctx.check(ctx.lib.sg_ball2d_fetch(ctx.h, flags, vp(q1h), vp(v1h), C.byref(c)))
        ctx.synchronize()
        t_e2e_local = time.perf_counter() - t0
        n_act, e2e_api = int(c.n_active), ("per rank: sg_ball2d_upload (owned q0,v0) + slab step (%s halo) + sg_ball2d_fetch (owned q1,v1 + this rank's contact list in global indices), pinned host buffers, wall clock; "
                                           "the per-rank lists are each ascending and merge into the reference's order body by body (sg_slab_merge_dest) -- not inside this timed region" % args.transport)
    h2d = 2 * 2 * n_local * 8
    d2h = 2 * 2 * n_local * 8 + n_act * (4 + 4 + 4 + 16 + 16 + 8)
    # the sampler ran from just before the timed steps to here (timed region, per-kernel pass, e2e pass: all under load)
    clocks = sampler.stop() if sampler else None

    # ---------------- reduce over ranks: max time, summed work ----------------
    if world > 1:
        tt = torch.tensor([t_local, t_e2e_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ww = torch.tensor([pairs_local, float(pc), float(pa), float(h2d), float(d2h), float(halo[0] + halo[1])], dtype=torch.float64, device="cuda")
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        t_max, t_e2e = float(tt[0]), float(tt[1])
        pairs_all, pc_all, pa_all, h2d_all, d2h_all, halo_all = [float(x) for x in ww]
    else:
        t_max, t_e2e, pairs_all, pc_all, pa_all, h2d_all, d2h_all, halo_all = t_local, t_e2e_local, pairs_local, float(pc), float(pa), float(h2d), float(d2h), 0.0

    # ---------------- N = 1 extra: the config round 1 was quoted on ----------------
    extra = None
    if world == 1 and args.config == 3 and not args.no_extras:
        del sim
        s2 = make_scene(2)
        sim2 = sb.Ball2DSim(sb.Ball2DState(s2["r"], s2["m"], s2["g"], s2["plane_x"], s2["plane_n"], s2["drum_x"], s2["drum_r"]), ctx=ctx)
        sim2.upload(s2["q"], s2["v"])
        for _ in range(3):
            ctx.flush_l2(); sim2.step(sb.SymplecticEulerMap(), s2["dt"])
        ms2 = []
        for _ in range(20):
            ctx.flush_l2(); ctx.timer_begin(); c2 = sim2.step(sb.SymplecticEulerMap(), s2["dt"]); ms2.append(ctx.timer_end())
        ctx.profile_enable(True); ctx.profile_reset()
        for _ in range(20):
            ctx.flush_l2(); sim2.step(sb.SymplecticEulerMap(), s2["dt"])
        p2 = ctx.profile(); ctx.profile_enable(False)
        extra = {"workload": workload_name(2), "value": (c2[0] + c2[1]) * len(ms2) / (sum(ms2) / 1e3), "unit": UNIT, "ms_per_step": sum(ms2) / len(ms2), "candidates": c2[0], "active": c2[1],
                 "kernels_us": {k: round(1e3 * v[1] / 20, 1) for k, v in sorted(p2.items(), key=lambda kv: -kv[1][1])},
                 "kernels_alg_GBps": {k: round(v[2] / (v[1] * 1e-3) / 1e9, 0) for k, v in p2.items() if v[1] > 0}}

    if rank == 0:
        peak, peak_src = measured_peak()
        total_ms = sum(v[1] for v in prof.values())
        top = max(prof.items(), key=lambda kv: kv[1][1])
        name, (nl, ms, by) = top
        achieved = (by / nl) / (ms / nl * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(name, args.config, n_local)
        kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps, "share": v[1] / total_ms,
                       "alg_GBps": (v[2] / (v[1] * 1e-3) / 1e9) if v[1] > 0 else None} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        cfg = config_dict(args.config, n_total)
        line = {
            "metric": METRIC, "value": pairs_all * args.steps / t_max, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "run": {"bodies_rank0": n_local, "candidates": pc_all, "active": pa_all, "candidate_pairs_per_s": pc_all * args.steps / t_max, "halo_bodies_per_step": halo_all,
                    "l2": "384 MB buffer overwritten before every timed step (L2 flush)", "timing": "CUDA events on the library stream, per step, summed; max over ranks",
                    "parallelism": ("%d equal-count x-quantile slabs of the randomly numbered scene, 1 process per GPU, ghost bodies exchanged with +-1 neighbours each step (%s), pair owned by the rank of its lower global index, "
                                    "per-step agreement on re-partitioning (one 4-byte all_reduce)" % (world, "peer-memory mailboxes over NVLink, no collective on the data path" if args.transport == "p2p" else "NCCL all_gather + send/recv")) if world > 1 else "single GPU"},
            "host_cores_bound": numa,
            "steps_per_s": args.steps / t_max,
            "step_ms_rank0": [round(x, 4) for x in step_ms],
            "clocks": clocks,
            "gpu_launches": gpu_launches,
            "parity_check": parity,
            "e2e": {"value": pairs_all * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all, "ms_per_step": 1e3 * t_e2e / e2e_steps,
                    "api": e2e_api},
            "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": by / nl,
                         "peak_source": peak_src, "share_of_step": ms / total_ms, "timed": "separate pass of the same %d steps with CUDA events around every kernel (rank 0)" % args.steps,
                         "kernels": kernels},
        }
        if e2e_shim is not None:
            line["e2e_indices_only"] = e2e_shim
        if extra is not None:
            line["config2"] = extra
        if not args.no_cpu_baseline and world == 1:   # the CPU figure is reported at N = 1 only (the other N of a scaling run would just repeat it)
            sc, sample = cpu_sample_scene(args.config)
            pairs_cpu, times, ckind, how = cpu_reference_step(sc, 2, 1)
            v = pairs_cpu * len(times) / sum(times)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": ckind,
                                    "sample": "%d timed steps of %s; %s; the reference's hot path is single-threaded" % (len(times), sample, how)}
            line["cpu_baseline_parallel"] = cpu_parallel_guarded(args.config, 3, 1)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
