#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/launch_summary.py gpurun_out/launches.csv out.csv "<command that was profiled>" """
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h, data = rows[0], rows[1:]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)
    agg.setdefault(r[ik][:70], []).append(v)
tot = sum(sum(v) for v in agg.values())
out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 400  {sys.argv[3] if len(sys.argv) > 3 else ''}",
       "# (cold-cache, serialised launches: compare SHARES; bench.py numbers come from the un-profiled run)",
       f"# total profiled device time {tot / 1e3:.3f} ms over {sum(len(v) for v in agg.values())} launches",
       "kernel,launches,total_us,share_pct,avg_us,min_us"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"\"{k}\",{len(v)},{sum(v):.1f},{100 * sum(v) / tot:.1f},{sum(v) / len(v):.1f},{min(v):.1f}")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[:8]))
