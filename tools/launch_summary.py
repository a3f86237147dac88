#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share.
usage: tools/launch_summary.py launches.csv [launches_per_step]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if per_step:
        rows = rows[-per_step:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("void cfm::<unnamed>::", "")[-80:]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot:.1f} us total (cold-cache, serialised under ncu: compare shares)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:9.1f} us {v[0]:4d}x {v[1] / v[0]:8.1f} us/launch {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
