#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep: SASS lines with the most warp-stall samples and their dominant reason.
usage: ncu_stalls.py REP [kernel-regex] [launch-skip] [top-n]"""
import csv, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:110])
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
print("total samples", tot)
agg = {}
for r in body:
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[idx[c]] or 0)
print("by reason:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
body.sort(key=lambda r: -int(r[idx["# Samples"]] or 0))
for r in body[:topn]:
    n = int(r[idx["# Samples"]] or 0)
    top = max(stall_cols, key=lambda c: int(r[idx[c]] or 0))
    print(f"{100*n/tot:5.1f}%  {r[idx['Address']][-5:]}  {top[6:]:14s} {r[idx['Source']][:90]}")
