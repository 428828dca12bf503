#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print(r[ki].split("(")[0])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"    {w:62s} {r[i]:>16s} {units[i]}")
