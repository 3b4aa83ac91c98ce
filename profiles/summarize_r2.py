#!/usr/bin/env python
"""gpurun_out/r2_{time,launches,full}_<tag>.* (profiles/capture_r2.sh) -> profiles/ncu_r2_<tag>.json / .md, one pair per config.
For every kernel: the LAST launch in the capture (a warm step), CUDA-event time from the un-profiled run beside ncu's own."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PRO = os.path.join(ROOT, "profiles")
PEAK = 6550.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

WANT = {"gpu__time_duration.sum": "ncu_time_us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
        "smsp__inst_executed.sum": "warp_insts", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_wavefront_pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst"}
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def short(name):
    return name.split("(")[0].replace("void ", "").strip()


def event_key(kernel):
    """kernel function name -> the name SG_LAUNCH times it under"""
    k = kernel
    table = [("sg_bp_count", "bp_count"), ("sg_bp_emit", "bp_emit"), ("sg_bp_contacts", "bp_contacts"), ("sg_bp_scatter", "bp_scatter"), ("sg_bp_hist", "bp_hist"),
             ("sg_bp_side_arrays", "bp_side"), ("sg_bp_bounds<", "bp_bounds"), ("k_ball2d_prep", "ball2d_flow_prep"), ("k_ball2d_static_emit", "ball2d_static_emit"),
             ("k_rb3d_flow", "rb3d_flow"), ("k_rb3d_sphere_static<(bool)0", "rb3d_plane_count"), ("k_rb3d_sphere_static<(bool)1", "rb3d_plane_emit"),
             ("k_rb3d_mesh_pairs<(bool)0", "rb3d_mesh_count"), ("k_rb3d_mesh_pairs<(bool)1", "rb3d_mesh_emit"), ("k_rb3d_aabb", "rb3d_aabb"),
             ("k_rb3d_pairs<(bool)0", "rb3d_pairs_count"), ("k_rb3d_pairs<(bool)1", "rb3d_pairs_emit")]
    for a, b in table:
        if a in k:
            return b
    return None


for tag in sys.argv[1:] or ["c2", "c3_2m", "c3_16m", "c4", "c5"]:
    tp = os.path.join(OUT, "r2_time_%s.json" % tag)
    if not os.path.exists(tp):
        continue
    timing = json.loads([l for l in open(tp).read().splitlines() if l.startswith("{")][-1])
    ev = timing["kernels"]
    # launch list: the last step = everything after the last flow kernel of the run
    launches = []
    lp = os.path.join(OUT, "r2_launches_%s.csv" % tag)
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
        names = [short(r[4]) for r in rows]
        first = max([i for i, n in enumerate(names) if "k_ball2d_prep" in n or "k_rb3d_flow" in n or "k_rb2d_flow" in n or "k_ball2d_flow" in n] or [0])
        step = rows[first:]
        tot = sum(float(r[-1]) for r in step)
        launches = [{"kernel": short(r[4]), "grid": r[8], "block": r[7], "ns": float(r[-1]), "share": float(r[-1]) / tot} for r in step]
    kern = {}
    fp = os.path.join(OUT, "r2_full_%s.csv" % tag)
    if os.path.exists(fp) and os.path.getsize(fp) > 0:
        rr = list(csv.reader(open(fp)))
        hdr, units = rr[0], rr[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rr[2:]:
            d = {}
            for m, k in WANT.items():
                if m in col and r[col[m]] not in ("", "no data", "n/a"):
                    d[k] = float(r[col[m]]) * (SCALE.get(units[col[m]], 1.0) if k.endswith("_MB") or k.endswith("_us") else 1.0)
            kern[short(r[col["Kernel Name"]])] = d   # later launches overwrite earlier ones: the warm step stays
    out = {"workload": timing["workload"], "bodies": timing["bodies"], "candidates": timing["candidates"], "active": timing["active"], "ms_per_step": timing["ms_per_step"],
           "pairs_per_s": timing["pairs_per_s"], "hbm_peak_GBps": PEAK, "event_kernels": ev, "ncu_kernels": kern, "launch_list_one_step": launches}
    json.dump(out, open(os.path.join(PRO, "ncu_r2_%s.json" % tag), "w"), indent=1)
    with open(os.path.join(PRO, "ncu_r2_%s.md" % tag), "w") as f:
        f.write("# Round 2, %s\n\n" % timing["workload"])
        f.write("%d bodies, %d candidate pairs, %d contacts; resident step %.4f ms = %.3g pairs/s (CUDA events, L2 flushed per step, not under a profiler).\n\n" % (
            timing["bodies"], timing["candidates"], timing["active"], timing["ms_per_step"], timing["pairs_per_s"]))
        f.write("| kernel (SG_LAUNCH name) | us/step (CUDA events) | share | algorithmic GB/s | of measured HBM peak %.0f GB/s |\n|---|---|---|---|---|\n" % PEAK)
        tot_us = sum(v["us"] for v in ev.values())
        for k, v in sorted(ev.items(), key=lambda kv: -kv[1]["us"]):
            f.write("| %s | %.1f | %.3f | %.0f | %.2f |\n" % (k, v["us"], v["us"] / tot_us, v["alg_GBps"], v["alg_GBps"] / PEAK))
        if kern:
            f.write("\n`ncu --set full --clock-control none`, last launch of each kernel in a warm step (times are cold-cache and serialised):\n\n")
            f.write("| kernel | ncu us | DRAM read MB | DRAM write MB | DRAM % | occupancy % | regs | warp insts | issue active % | L1 wavefront % | threads/inst | L1 hit % | L2 hit % | smem bank conflicts |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for k, d in sorted(kern.items(), key=lambda kv: -kv[1].get("ncu_time_us", 0)):
                g = lambda key, fmt="%.0f": (fmt % d[key]) if key in d else "-"
                f.write("| `%s` | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |\n" % (
                    k, g("ncu_time_us", "%.1f"), g("dram_read_MB", "%.1f"), g("dram_write_MB", "%.1f"), g("dram_pct"), g("occupancy_pct"), g("regs"), g("warp_insts", "%.3g"),
                    g("issue_active_pct"), g("l1_wavefront_pct"), g("threads_per_inst", "%.1f"), g("l1_hit_pct"), g("l2_hit_pct"), g("smem_bank_conflicts", "%.3g")))
        if launches:
            f.write("\nLaunch list of one step (`ncu --metrics gpu__time_duration.sum --clock-control none`; shares must agree with the event table, absolutes are cold-cache):\n\n| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|\n")
            for i, l in enumerate(launches):
                f.write("| %d | `%s` | %s | %s | %.1f | %.3f |\n" % (i + 1, l["kernel"], l["grid"], l["block"], l["ns"] / 1e3, l["share"]))
    print("wrote", tag)
