#!/usr/bin/env python
"""Top SASS instructions by stall samples with their dominant stall reasons: hot_sass.py file.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][0], 16)
recs = []
for r in rows[2:]:
    try:
        recs.append((int(r[iS]), int(r[iI]), r[iT], int(r[0], 16) - base, r[1].strip(), sorted(((int(r[i]), h[6:]) for i, h in stall_cols if r[i] not in ("", "0")), reverse=True)[:3]))
    except ValueError:
        pass
tot = sum(x[0] for x in recs) or 1
print(f"total samples {tot}, instructions {sum(x[1] for x in recs)}")
for s, n, t, off, ins, st in sorted(recs, reverse=True)[:top]:
    print(f"{100*s/tot:5.2f}% {n/1e3:8.0f}k thr={t:>5} +{off:05x} {ins[:70]:70s} {st}")
