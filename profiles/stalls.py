#!/usr/bin/env python
"""Top stall sites of an .ncu-rep (source page): python profiles/stalls.py file.ncu-rep [n]"""
import csv
import io
import subprocess
import sys

path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the report may hold several launches: split on "Kernel Name" records, use the first
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seg = rows[start[0] + 1: start[1] if len(start) > 1 else None]
hdr = seg[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for k, r in enumerate(seg[1:]):
    try:
        data.append((float(r[ix['Warp Stall Sampling (All Samples)']]), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
agg = {c: sum(float(d[2][ix[c]] or 0) for d in data) for c in stall_cols}
print('kernel', rows[start[0]][1][:80], ' samples', tot, ' instructions', len(data))
print('stall mix:', ', '.join('%s %.1f%%' % (c[6:], 100 * v / tot) for c, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for v, k, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = max(stall_cols, key=lambda c: float(r[ix[c]] or 0))
    print('%5.1f%% #%4d exec=%9s %-10s %s' % (100 * v / tot, k, r[ix['Instructions Executed']], top[6:], r[ix['Source']].strip()[:80]))
