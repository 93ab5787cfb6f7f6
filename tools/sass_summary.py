"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, bulk TMA -> UBLKCP / UTMALDG, tcgen05.commit / mbarrier -> UTCBAR / SYNCS) in libtdmpc_b200.so.
  python tools/sass_summary.py [lib.so] [out.txt]      (build() runs it and writes profiles/r2_sass_summary.txt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tdmpc_b200", "libtdmpc_b200.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "HMMA", "LDGSTS", "BAR.SYNC", "STL", "LDL"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, total, cur = collections.OrderedDict(), collections.Counter(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
lines = [f"SASS summary of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a): instructions and watched mnemonics per kernel",
         "tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS; HMMA would be the legacy mma.sync path",
         ""]
for k, c in counts.items():
    lines.append(f"{total[k]:7d} instr  {k[:110]}")
    if c:
        lines.append("               " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
lines += ["", "whole library: " + "  ".join(f"{w}={tot[w]}" for w in WATCH if tot[w])]
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-1:]), "->", os.path.relpath(out, ROOT))
