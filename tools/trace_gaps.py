"""Per-group epilogue gaps / waits / drains / math of a tools/trace_rollout.py dump.  python tools/trace_gaps.py dump.txt [first_job last_job]"""
import re, sys
fn = sys.argv[1]; lo = int(sys.argv[2]) if len(sys.argv) > 2 else 34; hi = int(sys.argv[3]) if len(sys.argv) > 3 else 56
rows = []
for line in open(fn):
    m = re.match(r'\s*(\d+) (\S+ \S+ c\d t\d)\s*\|\s*(\d+)\s+(\d+)\s+(\d+) \|\s*(\d+)\s+(\d+)\s+(\d+) \|\s*(\d+)\s+(\d+)\s+(\d+)(.*)', line)
    if not m: continue
    e = [(int(m.group(6)), int(m.group(7))), (int(m.group(9)), int(m.group(10)))]
    ext = re.findall(r'g(\d) wait\s+(-?\d+) drain\s+(-?\d+) math\s+(-?\d+)', m.group(12))
    rows.append((int(m.group(1)), m.group(2), int(m.group(3)), int(m.group(4)), e, ext))
last = [None, None]
for j, name, ms, me, e, ext in rows:
    if not (lo <= j <= hi): continue
    out = f"{j:3d} {name:14s} mma {ms:7d}-{me:7d} ({me - ms:5d})"
    for g in range(2):
        s_, e_ = e[g]
        x = [t for t in ext if int(t[0]) == g]
        if s_ and x:
            gap = s_ - last[g] if last[g] else 0
            out += f" | g{g} gap {gap:5d} start {s_:7d} wait {x[0][1]:>5s} drain {x[0][2]:>4s} math {x[0][3]:>5s} end {e_:7d}"
            last[g] = e_
    print(out)
