"""Summarise an `ncu --page source --csv` dump of rollout_tc_kernel: samples per mbarrier wait site
and top stalled instructions.  Usage: python tools/ncu_roles.py src.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = lambda i: int(data[i][ci['# Samples']])
tot = sum(S(i) for i in range(len(data)))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('total samples', tot, 'instructions', len(data))
d = {h: sum(int(r[ci[h]]) for r in data) for h in stalls}
print('stall mix:', [(k[6:], v) for k, v in sorted(d.items(), key=lambda x: -x[1])[:8]])
names = {0x0: 'w_full', 0x30: 'w_empty', 0x60: 'acc_full', 0x70: 'acc_empty', 0x80: 'h1_full', 0xa0: 'xs_full', 0xc0: 'za_full', 0xc8: 'q_full', 0xe8: '?'}
agg = {}
for i, r in enumerate(data):
    src = r[ci['Source']]
    if 'TRYWAIT' in src:
        m = re.search(r'\+0x([0-9a-f]+)\]', src)
        off = (int(m.group(1), 16) if m else 0) & 0xff
        key = names[max(k for k in names if k <= off)]
        s = sum(S(j) for j in range(i, min(i + 3, len(data))))
        agg.setdefault(key, [0, 0]); agg[key][0] += s
print('samples at mbarrier waits:', {k: v[0] for k, v in agg.items()})
top = sorted(range(len(data)), key=lambda i: -S(i))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for i in top:
    r = data[i]
    st = sorted(((int(r[ci[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{S(i):7d} {100*S(i)/tot:5.1f}% idx={i:5d} ex={r[ci['Instructions Executed']]:>9s} {r[ci['Source']].strip()[:64]:64s} {st}")
