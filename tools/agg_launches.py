#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/agg_launches.py gpurun_out/train_launches.csv [last_n]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
rows = []
for row in r:
    if len(row) <= vi:
        continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    rows.append((row[ki], v))
if last_n:
    rows = rows[-last_n:]
agg = collections.defaultdict(lambda: [0, 0.0])
for name, v in rows:
    nm = re.sub(r"\(.*", "", name)
    nm = re.sub(r"^void ", "", nm).split("::")[-1][:64]
    agg[nm][0] += 1
    agg[nm][1] += v
tot = sum(v for _, v in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us total")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v:10.1f} us {100 * v / tot:5.1f}% {c:5d}  {k}")
