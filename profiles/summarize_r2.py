#!/usr/bin/env python
"""Round 2: turn gpurun_out/prof_<name>_<tag>.ncu-rep into profiles/<tag>_<name>_ncu_summary.txt and
profiles/r2_traffic.json (DRAM bytes per launch of each shape's kernel, read by bench.py as roofline.traffic).
    python profiles/summarize_r2.py <tag>"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
ROWS = {10: 59_986_052, 100: 600_037_902}
# report name -> (traffic key, index of the shape's kernel inside the report, scale factor, bytes per row).  The Q3 reports
# hold four launches (customer, orders, lineitem scan = stages A + B, lineitem stage C): traffic and time of the lineitem
# side are the sums of the last two.
WHAT = {"q6": ("q6_sf100", 0, 100, 40), "q1": ("q1_sf100", 0, 100, 80), "q1d": ("q1d_sf10", 0, 10, 72),
        "q3_sf10": ("q3_sf10", 2, 10, 36), "q3_sf100": ("q3_sf100", 2, 100, 36)}
traffic_path = os.path.join(P, "r2_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for name, (key, idx, sf, bpr) in WHAT.items():
    rep = os.path.join(G, f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run([sys.executable, os.path.join(P, "read_ncu.py"), rep], capture_output=True, text=True).stdout
    hot = subprocess.run([sys.executable, os.path.join(P, "hot_lines.py"), rep, "24"], capture_output=True, text=True).stdout
    out_name = f"{tag}_{name}_ncu_summary.txt"
    with open(os.path.join(P, out_name), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; report gpurun_out/prof_{name}_{tag}.ncu-rep (scratch, not committed)\n")
        f.write(f"# command: python profiles/run_shape.py {name.split('_')[0]} {ROWS[sf]} 3   (see profiles/r2_capture.sh)\n")
        f.write(txt + "\n# source lines by stall samples (all launches of the report)\n" + hot)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    parts = [rows[2 + idx]]
    if name.startswith("q3") and len(rows) > 3 + idx and "entries_pipeline" in rows[3 + idx][hdr.index("Kernel Name")]:
        parts.append(rows[3 + idx])
    r = parts[0]
    def val(m):
        return sum(float(q[hdr.index(m)]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[hdr.index(m)]] for q in parts)
    dur = sum(float(q[hdr.index("gpu__time_duration.sum")]) for q in parts) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[units[hdr.index("gpu__time_duration.sum")]]
    traffic[key] = {"rows": ROWS[sf], "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                    "algorithmic_bytes": ROWS[sf] * bpr, "kernel_us_under_ncu": dur, "kernel": " + ".join(q[hdr.index("Kernel Name")][:60] for q in parts),
                    "source": f"profiles/{out_name} (ncu --set full --clock-control none, one launch)"}
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
for f in (f"launches_{tag}.csv", f"bench_{tag}.json", f"bench_{tag}_reference_arm.json", f"{tag}_tests.log"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f))
print(json.dumps({k: (round(v["dram_bytes_per_launch"] / 1e9, 3), round(v["algorithmic_bytes"] / 1e9, 3), v["kernel_us_under_ncu"]) for k, v in traffic.items()}))
