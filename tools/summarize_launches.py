#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py <launches.csv>"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
    k = r[ki].split("(")[0]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"{'kernel':62s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:62s} {c:8d} {v:10.2f} {100 * v / tot:6.1f}%")
print(f"{'total':62s} {sum(c for c, _ in agg.values()):8d} {tot:10.2f}")
