#!/usr/bin/env python
"""ncu launch list (gpu__time_duration.sum per launch, CSV) -> per-kernel totals and shares.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].split("::")[-1]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[14]) / 1e3          # ns -> us
tot = sum(a[1] for a in agg.values())
print(f"# {len(rows)} launches, {tot/1e3:.3f} ms of kernel time (ncu per-launch times are cold-cache and serialised:")
print("# compare SHARES with bench.py's kernels_ms_per_step, not absolute values)")
print(f"{'kernel':32s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'us/launch':>10s}")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:32s} {n:8d} {us:12.1f} {100*us/tot:6.1f}% {us/n:10.1f}")
