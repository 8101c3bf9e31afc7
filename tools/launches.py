#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum CSV): per-kernel totals and, with --eval N,
the per-launch list of the N-th velocity evaluation (delimited by time_embed_kernel launches)."""
import collections, csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
ev = int(sys.argv[2]) if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Unit")
def us(r):
    v = float(r[vi].replace(",", ""))
    return v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
data = [(r[ki].split("(")[0].split("::")[-1], us(r), r[gi]) for r in rows[1:]]
idx = [i for i, d in enumerate(data) if d[0].startswith("time_embed")]
if ev is None:
    tot, cnt = collections.defaultdict(float), collections.Counter()
    s, e = (idx[1], idx[2]) if len(idx) > 2 else (0, len(data))
    for n, v, g in data[s:e]:
        tot[n] += v; cnt[n] += 1
    print(f"one evaluation: {e - s} launches, {sum(tot.values()):.1f} us")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{v:10.1f} us {cnt[k]:4d}  {k}")
else:
    s, e = idx[ev], idx[ev + 1]
    for i in range(s, e):
        n, v, g = data[i]
        print(f"{i - s:4d} {v:9.1f} {g:>14s} {n[:40]}")
