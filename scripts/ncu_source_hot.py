#!/usr/bin/env python
"""Hottest SASS instructions (by warp-stall samples) from `ncu -i rep --page source --csv`, with stall reason columns."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = ci['# Samples']; src = ci['Source']; ex = ci['Instructions Executed']
reasons = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[S] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
top = sorted(data, key=lambda r: -int(r[S] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for r in top:
    n = int(r[S] or 0)
    rs = sorted(((int(r[ci[k]] or 0), k[6:]) for k in reasons), reverse=True)[:3]
    print(f'{100*n/tot:5.1f}% {n:7d} exec={r[ex]:>9s}  {r[src][:70]:70s} ' + ' '.join(f'{k}:{v}' for v, k in rs if v))
# region split: samples by opcode class
import collections
cls = collections.Counter()
for r in data:
    op = r[src].split()[0] if r[src] else ''
    if op.startswith('@'):
        op = r[src].split()[1]
    cls[op.split('.')[0]] += int(r[S] or 0)
print('by opcode:', ', '.join(f'{k}:{100*v/tot:.1f}%' for k, v in cls.most_common(14)))
