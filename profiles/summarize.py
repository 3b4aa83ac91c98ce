#!/usr/bin/env python
"""Turns gpurun_out/{launches_RR.csv, prof_RR.ncu-rep, bench_RR_1gpu.json} into the tracked summaries
profiles/launches_RR.csv (one step), profiles/ncu_RR.json and profiles/ncu_RR.md.   python profiles/summarize.py r1"""
import csv
import json
import os
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PRO = os.path.join(ROOT, "profiles")

# ---- launch list: keep the last full step (from one flow kernel to the next) ----
rows = [r for r in csv.reader(open(os.path.join(OUT, "launches_%s.csv" % R))) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]
starts = [i for i, n in enumerate(names) if "k_ball2d_prep<1>" in n.replace("(bool)1", "1") or "k_ball2d_prep<true>" in n or ("k_ball2d_prep" in n and "true" in n)]
if len(starts) < 2:
    starts = [i for i, n in enumerate(names) if "k_ball2d_prep" in n]
s, e = starts[-2], starts[-1]
step = rows[s:e]
tot = sum(float(r[-1]) for r in step)
with open(os.path.join(PRO, "launches_%s.csv" % R), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 3 (one step = %d launches, %.1f us of kernel time; cold-cache, serialised)\n" % (len(step), tot / 1e3))
    f.write("kernel,grid,block,ns,share\n")
    for r in step:
        f.write('"%s","%s","%s",%s,%.4f\n' % (r[4].split("(")[0].replace("void ", ""), r[8], r[7], r[-1], float(r[-1]) / tot))

# ---- full capture ----
raw = subprocess.run(["ncu", "-i", os.path.join(OUT, "prof_%s.ncu-rep" % R), "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
col = {h: i for i, h in enumerate(hdr)}
want = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
        "smsp__inst_executed.sum": "warp_insts", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefront_pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct"}


def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


kern = {}
for r in rr[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    d = {}
    for m, k in want.items():
        if m in col and r[col[m]] not in ("", "no data"):
            d[k] = to_bytes(r[col[m]], units[col[m]]) if k.startswith("dram_") and k != "dram_pct" else float(r[col[m]])
    if "time_us" in d and units[col["gpu__time_duration.sum"]] == "ms":
        d["time_us"] *= 1e3
    if "time_us" in d and units[col["gpu__time_duration.sum"]] == "ns":
        d["time_us"] /= 1e3
    d["dram_traffic"] = d.get("dram_read", 0) + d.get("dram_write", 0)
    kern[name] = d
json.dump(kern, open(os.path.join(PRO, "ncu_%s.json" % R), "w"), indent=1)

bench = None
bp = os.path.join(OUT, "bench_%s_1gpu.json" % R)
if os.path.exists(bp):
    txt = [l for l in open(bp).read().splitlines() if l.startswith("{")]
    if txt:
        bench = json.loads(txt[-1])
        json.dump(bench, open(os.path.join(PRO, "bench_%s_1gpu.json" % R), "w"), indent=1)
rp = os.path.join(OUT, "bench_%s_reference.json" % R)
if os.path.exists(rp):
    txt = [l for l in open(rp).read().splitlines() if l.startswith("{")]
    if txt:
        json.dump(json.loads(txt[-1]), open(os.path.join(PRO, "bench_%s_reference.json" % R), "w"), indent=1)

with open(os.path.join(PRO, "ncu_%s.md" % R), "w") as f:
    f.write("# ncu `--set full --clock-control none` capture, round %s (config 2: 1M balls)\n\n" % R)
    f.write("| kernel | time us | DRAM read MB | DRAM write MB | occupancy % | regs | warp insts | issue active % | LSU wavefront % | smem bank conflicts | L2 hit % |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for k, d in kern.items():
        f.write("| `%s` | %.1f | %.1f | %.1f | %.0f | %.0f | %.3g | %.0f | %.0f | %.3g | %.0f |\n" % (
            k, d.get("time_us", 0), d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6, d.get("occupancy_pct", 0), d.get("regs", 0), d.get("warp_insts", 0),
            d.get("issue_active_pct", 0), d.get("lsu_wavefront_pct", 0), d.get("smem_bank_conflicts", 0), d.get("l2_hit_pct", 0)))
    if bench:
        f.write("\nbench line of the same build: value %.4g %s, %.4f ms/step, e2e %.4g (%.3f ms/step), clocks %s\n\n" % (
            bench["value"], bench["unit"], bench["ms_per_step"], bench["e2e"]["value"], bench["e2e"]["ms_per_step"], json.dumps(bench.get("clocks"))))
        f.write("| kernel (CUDA events inside bench.py) | us/step | share | algorithmic GB/s |\n|---|---|---|---|\n")
        for k, v in bench["roofline"]["kernels"].items():
            f.write("| %s | %.1f | %.3f | %s |\n" % (k, v["ms_per_step"] * 1e3, v["share"], "%.0f" % v["alg_GBps"] if v["alg_GBps"] else "-"))
print("wrote profiles/launches_%s.csv, ncu_%s.json, ncu_%s.md" % (R, R, R))
