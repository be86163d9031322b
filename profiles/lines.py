#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA SOURCE LINE of one kernel of an .ncu-rep.
ncu's CSV source page is SASS only, so the line of every instruction comes from `nvdisasm -g` of the cubin
inside the library that ran (matched by code offset).

    python profiles/lines.py file.ncu-rep <kernel regex> <module, e.g. rle_flat> [n lines]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, kname, module = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lib = os.path.join(ROOT, 'ampis_b200', 'libampis_b200.so')

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', module + '.sm_100a.cubin', lib], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, module + '.sm_100a.cubin')
sass = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# per function: list of (offset, line)
funcs, cur, line = {}, None, None
for l in sass.splitlines():
    m = re.match(r'\s*\.global\s+(\S+)', l)
    if m:
        cur = m.group(1); funcs[cur] = {}; line = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);', l)
    if m and cur:
        funcs[cur][int(m.group(1), 16)] = line

txt = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
start = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
seg = rows[start[0] + 1: start[1] if len(start) > 1 else None]
kfull = rows[start[0]][1]
hdr = seg[0]
ix = {h: i for i, h in enumerate(hdr)}
base = int(seg[1][ix['Address']], 16)
short = kfull.split('(')[0].split()[-1].split('<')[0]
cands = [f for f in funcs if short in f]
lines_of = funcs[cands[0]] if cands else {}
agg = collections.OrderedDict()
tot_i = tot_s = 0.0
for r in seg[1:]:
    try:
        ie = float(r[ix['Instructions Executed']] or 0)
        ss = float(r[ix['Warp Stall Sampling (All Samples)']] or 0)
    except Exception:
        continue
    ln = lines_of.get(int(r[ix['Address']], 16) - base)
    a = agg.setdefault(ln, [0.0, 0.0, 0])
    a[0] += ie; a[1] += ss; a[2] += 1
    tot_i += ie; tot_s += ss
src_cache = {}
def text(ln):
    if ln is None:
        return '?'
    f = [p for p in (os.path.join(ROOT, 'ampis_b200', 'csrc', ln[0]),) if os.path.exists(p)]
    if not f:
        return ln[0]
    if f[0] not in src_cache:
        src_cache[f[0]] = open(f[0]).read().splitlines()
    L = src_cache[f[0]]
    return L[ln[1] - 1].strip()[:95] if ln[1] - 1 < len(L) else ''
print('kernel %s: %.0f warp instructions executed, %.0f stall samples, %d SASS instructions' % (short, tot_i, tot_s, len(seg) - 1))
for ln, (ie, ss, c) in sorted(agg.items(), key=lambda x: -x[1][0])[:n]:
    print('%5.1f%% inst %5.1f%% stall %4d sass  %s:%s  %s' % (100 * ie / tot_i, 100 * ss / max(tot_s, 1), c,
                                                             ln[0] if ln else '?', ln[1] if ln else '', text(ln)))
