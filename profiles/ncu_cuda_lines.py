#!/usr/bin/env python
"""Hot CUDA-C source lines from `ncu -i X.ncu-rep --page source --print-source cuda --csv --kernel-name regex:K`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = []
fname, hdr, col = None, None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        hdr = None
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        col = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        samples = int(r[col["# Samples"]] or 0)
        insts = int(r[col["Instructions Executed"]] or 0)
    except Exception:
        continue
    out.append((samples, insts, fname, r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out)
print("total samples", tot, "warp insts", sum(o[1] for o in out))
for s, i, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print("%5.1f%% %9d  %s:%s  %s" % (100.0 * s / max(tot, 1), i, f, ln, src))
