#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lines = [l for l in open(path) if not l.startswith('==')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
n = 0
for row in r:
    if len(row) <= vi:
        continue
    try:
        v = float(row[vi].replace(',', ''))
    except ValueError:
        continue
    k = re.sub(r'\(.*', '', row[ki])
    k = re.sub(r'<.*', '', k)[:70]
    agg[k][0] += 1; agg[k][1] += v; n += 1
tot = sum(v for _, v in agg.values())
print(f'{n} launches over {steps} step(s): {n/steps:.0f} launches/step, {tot/1e6/steps:.3f} ms/step (serialised, cold-cache kernel time)')
print(f'{"ms/step":>9s} {"share":>6s} {"n/step":>7s}  kernel')
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'{v/1e6/steps:9.3f} {100*v/tot:5.1f}% {c/steps:7.1f}  {k}')
