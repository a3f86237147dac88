#!/usr/bin/env python
"""Top stall sites from `ncu -i rep --page source --csv` output (all kernels concatenated).
usage: ncu_stalls.py src.csv [section_index] [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lo = starts[sec]; hi = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
print(len(starts), 'kernels;', rows[lo][1][:100])
hdr = rows[lo + 1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:topn]:
    s = {h[6:]: int(r[idx[h]]) for h in stalls if int(r[idx[h]]) > 0}
    print(r[idx['# Samples']].rjust(6), r[idx['Source']].strip()[:60].ljust(60), s)
