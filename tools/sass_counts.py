#!/usr/bin/env python
"""SASS instruction counts per kernel of libcfm_b200.so (cuobjdump -sass): which kernels carry tcgen05 / TMEM / TMA code.
usage: python tools/sass_counts.py > profiles/rN_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "conformer_pytorch_lightning_b200", "libcfm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|HMMA|MUFU\.EX2|MUFU\.TANH)\b")
counts = collections.defaultdict(collections.Counter)
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for t in pat.findall(line):
            counts[cur][t] += 1
names = list(counts)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
def _short(d):
    m = re.search(r"(\w+_kernel)", d)
    return m.group(1) if m else d[:64]


short = {n: _short(d) for n, d in zip(names, dem)}
keys = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU.EX2", "MUFU.TANH"]
print("SASS instruction counts per kernel of libcfm_b200.so (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2),")
print("LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, HMMA = legacy mma.sync, MUFU.* = SFU.\n")
print(f"{'kernel':64s} " + " ".join(f"{k:>12s}" for k in keys))
agg = collections.defaultdict(collections.Counter)
for fn, c in counts.items():
    agg[short.get(fn, fn)].update(c)                                # merge template instantiations
tot = collections.Counter()
for fn, c in sorted(agg.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCHMMA.2CTA"] + kv[1]["UTMALDG"])):
    if not any(c[k] for k in keys):
        continue
    print(f"{fn:64s} " + " ".join(f"{c[k]:12d}" for k in keys))
    tot.update(c)
print(f"{'TOTAL (all instantiations)':64s} " + " ".join(f"{tot[k]:12d}" for k in keys))
