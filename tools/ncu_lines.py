#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample totals of one kernel from an
.ncu-rep (needs -lineinfo and --import-source on).  Dev tool.
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k4_pairs [top]"""
import csv, io, subprocess, sys, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
agg = collections.OrderedDict()
srcs = {}
seen_launch = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= iT:
        continue
    if r[0].strip().isdigit() and r[2] == "-" and r[iI].isdigit():      # a CUDA line with its totals
        key = (cur_file, int(r[0]))
        srcs[key] = r[1].strip()
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += int(r[iS] or 0); a[1] += int(r[iI]); a[2] += int(r[iT])
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"{kern}: samples {tot_s}  warp-instr {tot_i}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[1]/tot_i:5.1f}% instr {100*a[0]/tot_s:5.1f}% samples  thr/instr {a[2]/max(a[1],1):4.1f}  {f}:{l}  {srcs[(f,l)][:90]}")
