#!/usr/bin/env python
"""Per-phase (BAR.SYNC-delimited) summary of an ncu source-page CSV: instructions/key, stall samples, smem wavefronts.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_phases.py src.csv <log2 keys> [min executions]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
WK = (1 << int(sys.argv[2])) / 32
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 2e5
ia = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); isamp = hdr.index("# Samples")
iws = hdr.index("L1 Wavefronts Shared"); iwi = hdr.index("L1 Wavefronts Shared Ideal")
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
def new(): return {'n': 0, 's': 0, 'ws': 0, 'wi': 0, 'st': [0] * len(stalls), 'first': None, 'cnt': 0, 'ops': {}}
seg = []; cur = new()
for idx, r in enumerate(rows[2:]):
    n = int(r[ia])
    if n < thr: continue
    if cur['first'] is None: cur['first'] = idx
    src = r[isrc].split()
    op = (src[1] if src[0].startswith('@') else src[0]).split('.')[0]
    cur['ops'][op] = cur['ops'].get(op, 0) + n
    cur['n'] += n; cur['s'] += int(r[isamp]); cur['ws'] += int(r[iws] or 0); cur['wi'] += int(r[iwi] or 0); cur['cnt'] += 1
    for k, si in enumerate(stalls): cur['st'][k] += int(r[si] or 0)
    if 'BAR.SYNC' in r[isrc]:
        seg.append(cur); cur = new()
seg.append(cur)
tot = sum(s['s'] for s in seg) or 1
for s in seg:
    top = sorted(zip(s['st'], [hdr[i][6:] for i in stalls]), reverse=True)[:4]
    ops = sorted(s['ops'].items(), key=lambda kv: -kv[1])[:7]
    print(f"line={s['first']:5d} sass={s['cnt']:4d} inst/key={s['n']/WK:6.2f} samples={100*s['s']/tot:5.1f}% "
          f"smem wf/key={s['ws']/WK:5.2f} (ideal {s['wi']/WK:5.2f})  stalls:", [(b, a) for a, b in top])
    print("        ", " ".join(f"{k}={v/WK:.2f}" for k, v in ops))
