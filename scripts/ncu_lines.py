#!/usr/bin/env python
"""Per CUDA source line warp-stall samples from `ncu -i rep --page source --print-source cuda,sass --csv`."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = collections.Counter(); text = {}; execs = collections.Counter()
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Line No':
        hdr = r; S = hdr.index('# Samples'); E = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) <= S: continue
    if r[0] and r[0].isdigit():          # a CUDA source line row (aggregated over its SASS)
        key = (cur_file, int(r[0]))
        agg[key] += int(r[S]) if r[S].isdigit() else 0; text[key] = r[1].strip()[:110]; execs[key] += int(r[E]) if r[E].isdigit() else 0
tot = sum(agg.values())
print('total samples', tot)
for (f, ln), n in agg.most_common(topn):
    print(f'{100*n/tot:5.1f}% {n:7d} ex={execs[(f,ln)]:>10d} {f}:{ln:<4d} {text[(f,ln)]}')
