#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep: python profiles/rawsum.py file.ncu-rep"""
import csv
import subprocess
import sys
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
want = [('Kernel Name', 'kernel'), ('gpu__time_duration.sum', 'us'), ('smsp__inst_executed.sum', 'inst'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long_sb'),
        ('l1tex__t_sector_hit_rate.pct', 'l1hit'), ('lts__t_sector_hit_rate.pct', 'l2hit')]
print(' | '.join(n for _, n in want), '   units:', [rows[1][ix[w]] for w, _ in want[1:7]])
for r in rows[2:]:
    out = []
    for w, n in want:
        v = r[ix[w]] if w in ix else ''
        if n == 'kernel':
            v = v.split('(')[0][-34:]
        else:
            try:
                v = '%.4g' % float(v.replace(',', ''))
            except ValueError:
                pass
        out.append(v)
    print(' | '.join(out))
