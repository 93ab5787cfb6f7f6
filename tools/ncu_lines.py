"""Per-source-line roll-up of an `ncu --page source --print-source cuda,sass --csv` dump: samples, executed warp
instructions and the top stall reasons per CUDA source line (file:line).  Usage: python tools/ncu_lines.py dump.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No":
        hdr = r; ci = {}
        for i, h in enumerate(hdr): ci.setdefault(h, i)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) != len(hdr): continue
    if r[ci["Address"]] != "-":            # SASS rows below a source line: skip (their totals are on the source row)
        continue
    try: s = int(r[ci["# Samples"]]); ex = int(r[ci["Instructions Executed"]])
    except ValueError: continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, {}, r[1].strip()[:90]])
    a[0] += s; a[1] += ex
    for h in stalls:
        v = int(r[ci[h]] or 0)
        if v: a[2][h[6:]] = a[2].get(h[6:], 0) + v
tot = sum(a[0] for a in agg.values()); totex = sum(a[1] for a in agg.values())
print("total samples", tot, "warp instructions", totex)
for key, a in sorted(agg.items(), key=lambda x: -x[1][0])[:N]:
    st = sorted(a[2].items(), key=lambda x: -x[1])[:3]
    print(f"{a[0]:7d} {100*a[0]/tot:5.1f}% ex={a[1]:>10d} {100*a[1]/totex:4.1f}% {key[0]}:{key[1]:<5d} {a[3]:90s} {st}")
