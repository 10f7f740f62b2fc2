#!/usr/bin/env python
"""Turn the ncu reports of one capture (gpurun_out/prof_*_<tag>.ncu-rep) into the committed
summaries under profiles/:  python profiles/summarize.py <tag>"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
for name in ("q6", "q1", "q3", "q3bloom", "bloom"):
    rep = os.path.join(G, f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run([sys.executable, os.path.join(P, "read_ncu.py"), rep], capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(P, "hot_lines.py"), rep, "15"], capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{tag}_{name}_ncu_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch; report gpurun_out/prof_{name}_{tag}.ncu-rep (scratch)\n")
        f.write(txt + "\n# source lines by stall samples\n" + hot)
    if name == "q6":
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units, r = rows[0], rows[1], rows[2]
        def val(m):
            v, u = float(r[hdr.index(m)]), units[hdr.index(m)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tj = {"rows": 59986052, "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
              "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
              "algorithmic_bytes": 59986052 * 40, "kernel_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")]),
              "source": f"profiles/{tag}_q6_ncu_summary.txt (ncu --set full --clock-control none, TPC-H Q6 shape SF10, one launch)"}
        json.dump(tj, open(os.path.join(P, "q6_sf10_traffic.json"), "w"), indent=1)
for f in (f"launches_{tag}.csv", f"bench_{tag}.json"):
    src = os.path.join(G, f)
    if os.path.exists(src):
        open(os.path.join(P, f), "w").write(open(src).read())
print("summaries written to profiles/")
