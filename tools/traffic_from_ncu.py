"""profiles/rollout_tc_traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log of
`tools/prof_rollout.py --batch B`: DRAM bytes per rollout_tc_kernel launch (mean over the captured launches).
  python tools/traffic_from_ncu.py log.csv workload batch"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log, wl, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = [r for r in csv.reader(l for l in open(log) if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = {}
for r in rows[1:]:
    if "rollout_tc" not in r[ki] or not r[mi].startswith("dram__bytes"):
        continue
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    per.setdefault(r[hdr.index("ID")], 0.0)
    per[r[hdr.index("ID")]] += float(r[vi].replace(",", "")) * scale
vals = list(per.values())
p = os.path.join(ROOT, "profiles", "rollout_tc_traffic.json")
d = json.load(open(p)) if os.path.exists(p) else {}
d[wl] = {"batch": batch, "dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
         "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum per rollout_tc_kernel launch at {batch} environments (tools/prof_rollout.py --batch {batch})"}
json.dump(d, open(p, "w"), indent=1)
print(d[wl])
