#!/usr/bin/env python
"""ncu launch list with DRAM bytes (CSV of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
    python tools/probe.py c3_human 0 1)
-> per-kernel DRAM bytes of ONE pipeline step (the last one in the file), under the names of the
library's own per-kernel timers, as profiles/rNN_ncu_traffic.json (bench.py reads it).
    python tools/ncu_traffic.py launches.csv c3_human 10000000 > profiles/r02_ncu_traffic.json"""
import collections
import csv
import json
import re
import sys

ALIAS = {"k2_partition2": "k2_partition", "k2_deliver2": "k2_deliver", "k4_pairs3": "k4_pairs",
         "k4_finalize2": "k4_finalize", "k_fire_rounds_all": "k_fire_round", "k4_fire_redo": "k4_fire_init",
         "k2_init_cursors": "k2_partition", "k2_init_group_cursors": "k2_deliver",
         "k_scan_tile_sums": "scan", "k_scan_tile_offsets": "scan", "k_scan_tiles": "scan",
         "k_poly_reset": "k_poly_sweep", "k_poly_propose": "k_poly_sweep", "k_poly_commit": "k_poly_sweep",
         "k2_lineless_count": "k2_lineless", "k2_lineless_assign": "k2_lineless", "k2_fill_ls": "k2_lineless"}

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
H = rows[hdr]
iK, iM, iV, iU, iID = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= iV or not r[iID].isdigit():
        continue
    name = re.sub(r"<.*", "", r[iK].split("(")[0].split("::")[-1]).strip()
    d = launch.setdefault(int(r[iID]), {"name": name})
    v = float(r[iV].replace(",", ""))
    u = r[iU]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0}.get(u, 1)
    d[r[iM]] = v * scale
ids = sorted(launch)
starts = [i for i in ids if launch[i]["name"] in ("k3_lines", "k2_head_counts")]
last = starts[-1]
prev_len = (starts[-1] - starts[-2]) if len(starts) > 1 else None
step = [launch[i] for i in ids if i >= last]
per = collections.OrderedDict()
for d in step:
    n = ALIAS.get(d["name"], d["name"])
    a = per.setdefault(n, {"launches": 0, "dram_bytes": 0.0, "ms_under_ncu": 0.0})
    a["launches"] += 1
    a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["ms_under_ncu"] += d.get("gpu__time_duration.sum", 0.0)
total = sum(a["dram_bytes"] for a in per.values())
out = {"workload": sys.argv[2], "vertices": int(sys.argv[3]),
       "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 "python tools/probe.py %s 0 1: the launches of the last pipeline step in the capture (%d launches%s); "
                 "bytes = dram read + write, summed over the launches of a kernel in that step"
                 % (sys.argv[2], len(step), "" if prev_len is None else ", the step before had %d" % prev_len),
       "dram_bytes_per_step": int(total),
       "dram_bytes_per_launch": {k: int(v["dram_bytes"]) for k, v in sorted(per.items(), key=lambda kv: -kv[1]["dram_bytes"])},
       "launches": {k: v["launches"] for k, v in per.items()},
       "ms_under_ncu": {k: round(v["ms_under_ncu"], 4) for k, v in per.items()}}
print(json.dumps(out, indent=1))
