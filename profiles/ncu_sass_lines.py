#!/usr/bin/env python
"""Join an ncu SASS-level source page with the line table of the SAME build, so that executed instructions and stall samples
can be read per CUDA source line (the ncu CLI in this image prints no metrics on its CUDA view):

  cuobjdump -xelf all scisim_b200/libscisim_b200.so          # -> sg_ball2d.sm_100a.cubin ...
  nvdisasm -g -c sg_ball2d.sm_100a.cubin > ball2d.sass
  python profiles/ncu_sass_lines.py gpurun_out/ncu_X_source.csv ball2d.sass '<mangled kernel name substring>' [top]

Instructions are matched by their offset from the kernel's first instruction."""
import csv
import re
import sys
from collections import defaultdict

csv_path, sass_path, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- line table of the kernel
line_of = {}
cur = None
inside = False
for l in open(sass_path):
    if l.startswith(".text."):
        inside = kname in l
        cur = None
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] != "Address"]
base = int(body[0][0], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot_i = tot_s = 0
miss = 0
for r in body:
    off = int(r[0], 16) - base
    ent = line_of.get(off)
    if ent is None:
        miss += 1
        key = ("?", 0)
    else:
        key = ent[0] or ("?", 0)
    insts = int(r[col["Instructions Executed"]] or 0)
    samp = int(r[col["# Samples"]] or 0)
    a = agg[key]
    a[0] += insts
    a[1] += samp
    for s in stalls:
        v = int(r[col[s]] or 0)
        if v:
            a[2][s[6:]] += v
    tot_i += insts
    tot_s += samp
print("kernel lines joined: %d SASS instructions, %d unmatched; warp insts %d, samples %d" % (len(body), miss, tot_i, tot_s))
src_cache = {}


def src(f, n):
    if f not in src_cache:
        try:
            import glob
            p = glob.glob("/root/repo/scisim_b200/csrc/" + f) or glob.glob("/usr/local/cuda/targets/x86_64-linux/include/**/" + f, recursive=True)
            src_cache[f] = open(p[0]).read().split("\n") if p else []
        except Exception:
            src_cache[f] = []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""


print("%6s %6s  %-28s %-22s %s" % ("inst%", "samp%", "file:line", "top stalls", "source"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join("%s=%d" % (k, v) for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:2])
    print("%6.2f %6.2f  %-28s %-22s %s" % (100.0 * a[0] / max(tot_i, 1), 100.0 * a[1] / max(tot_s, 1), "%s:%d" % key, st, src(*key)))
