#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of an ncu report captured with
--import-source on:  hot_lines.py file.ncu-rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
fname, hdr, cur = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]), r[1].strip()); agg.setdefault(cur, [0, 0, 0]); continue
    if cur is None:
        continue
    try:
        agg[cur][0] += int(r[iI]); agg[cur][1] += int(r[iS]); agg[cur][2] += 1
    except (ValueError, IndexError):
        pass
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
print(f"{'inst%':>6} {'samp%':>6} {'sass':>5}  file:line  source")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[0]/tot_i:6.2f} {100*v[1]/tot_s:6.2f} {v[2]:5d}  {k[0]}:{k[1]}  {k[2][:110]}")
