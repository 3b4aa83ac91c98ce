#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` : stall totals and the hottest SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
body = [r for r in rows[2:] if len(r) >= len(hdr) and r[0] != "Address"]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
samples = 0
for r in body:
    samples += int(r[col["# Samples"]] or 0)
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
print("kernel:", rows[0][1][:100])
print("SASS lines:", len(body), "samples:", samples, "warp insts:", sum(int(r[col["Instructions Executed"]] or 0) for r in body))
print("stalls:", ", ".join("%s=%.1f%%" % (s[6:], 100.0 * v / max(samples, 1)) for s, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.01 * samples))
print("hottest lines (samples, insts executed, sass, dominant stall):")
for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
    dom = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print("%6s %9s  %-70s %s" % (r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]].strip()[:70], dom[6:]))
