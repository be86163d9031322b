#!/usr/bin/env python
"""Turn the round-2 evidence brought back in gpurun_out/ (profiles/run_r02_evidence.sh <tag>) into tracked files:

    python profiles/summarize_r02.py r02j

  profiles/kernels_r02.md          ncu --set full of every kernel of the crop-layout step (1,000 and 91 images per
                                   launch) and of the string decode kernel of the e2e path
  profiles/launches_r02_crop.md    ncu launch list of one bench command: per-kernel device time and shares
  profiles/traffic.json            DRAM bytes per launch of the hot kernels (read by bench.py for roofline.traffic)
  profiles/bench_r02_*.json        the bench lines themselves
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02j'

WANT = [('gpu__time_duration.sum', 'time us'), ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
        ('dram__bytes_read.sum', 'DRAM read MB'), ('dram__bytes_write.sum', 'DRAM write MB'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
        ('launch__registers_per_thread', 'registers'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long-scoreboard stall / issue'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'), ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %')]
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3,
         'msecond': 1e3, 'second': 1e6, 'nsecond': 1e-3}


def raw(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {'kernel': r[ix['Kernel Name']].split('(')[0].replace('void ', '')}
        for m, name in WANT:
            if m not in ix:
                continue
            try:
                v = float(r[ix[m]].replace(',', ''))
            except ValueError:
                continue
            u = units[ix[m]]
            if 'time' in name or 'MB' in name:
                v *= SCALE.get(u, 1.0)
            d[name] = v
        out.append(d)
    return out


def table(rows):
    names = [n for _, n in WANT if any(n in r for r in rows)]
    md = ['| metric | ' + ' | '.join('`%s`' % r['kernel'] for r in rows) + ' |', '|---|' + '---|' * len(rows)]
    for n in names:
        md.append('| %s | ' % n + ' | '.join(('%.4g' % r[n]) if n in r else '' for r in rows) + ' |')
    return md


def main():
    md = ['# ncu --set full summaries, round 2 (%s)' % tag, '',
          'Commands: `profiles/run_r02_evidence.sh` (each capture follows a plain run of the same command that exited 0). '
          'C2 = c2_powder_batch, 1024x1024 frames, 500 x 500 masks per image, crop layout, dense matrices out. Times '
          'under ncu are cold-cache and serialised: compare shares and per-kernel counters, not absolute times with '
          '`bench.py`.', '']
    traffic = json.load(open(os.path.join(PROF, 'traffic.json'))) if os.path.exists(os.path.join(PROF, 'traffic.json')) else {}
    for name, n_img, title in (('crop1000_%s' % tag, 1000, '1,000 images per launch (what `bench.py` runs)'),
                               ('crop91_%s' % tag, 91, '91 images per launch (comparable with `kernels_r01p.md`)'),
                               ('strdec_%s' % tag, 250, 'e2e path: `rle_string_decode_kernel`, 250 images per call')):
        p = os.path.join(OUT, name + '.ncu-rep')
        if not os.path.exists(p):
            continue
        rows = raw(p)
        md += ['## ' + title, ''] + table(rows) + ['']
        for r in rows:
            tot = r.get('DRAM read MB', 0) + r.get('DRAM write MB', 0)
            md.append('* `%s`: %.1f MB of DRAM traffic in %.1f us = %.0f GB/s; %.0f warp instructions per mask'
                      % (r['kernel'], tot, r['time us'], tot / max(r['time us'], 1e-9) * 1e3,
                         r.get('warp instructions', 0) / (n_img * 1000.0)))
            if n_img == 1000 and r['kernel'].startswith('rle_flat'):
                traffic['c2_powder_batch/crop/paint'] = {'bytes_per_image': tot * 1e6 / n_img, 'bytes_per_launch': tot * 1e6,
                                                         'images_per_launch': n_img, 'source': 'profiles/kernels_r02.md'}
        if n_img == 1000:
            rows_k = [r for r in rows if r['kernel'].split('<')[0] in ('grid_build_kernel', 'pairs_from_grid_kernel',
                                                                      'pair_intersect_kernel', 'pair_intersect_flat_kernel',
                                                                      'rows_from_pairs_kernel')]
            tot = sum(r.get('DRAM read MB', 0) + r.get('DRAM write MB', 0) for r in rows_k)
            traffic['c2_powder_batch/crop/rows'] = {'bytes_per_image': tot * 1e6 / n_img, 'bytes_per_launch': tot * 1e6,
                                                    'images_per_launch': n_img, 'source': 'profiles/kernels_r02.md (grid build + the '
                                                    'three join kernels; the dense matrices are cleared by the decode kernel)'}
        md.append('')
    open(os.path.join(PROF, 'kernels_r02.md'), 'w').write('\n'.join(md) + '\n')
    json.dump(traffic, open(os.path.join(PROF, 'traffic.json'), 'w'), indent=1, sort_keys=True)

    # launch list
    p = os.path.join(OUT, 'launches_%s_crop1000.csv' % tag)
    if os.path.exists(p):
        lines = open(p).read().splitlines()
        for i, l in enumerate(lines):
            if l.startswith('"ID"'):
                rows = list(csv.DictReader(io.StringIO('\n'.join(lines[i:]))))
                break
        agg = collections.OrderedDict()
        for r in rows:
            k = r['Kernel Name'].split('(')[0].replace('void ', '')
            agg.setdefault(k, []).append(float(r['Metric Value']) / 1e3)
        step_prefixes = ('rle_flat_crop_kernel', 'rle_measure_paint_list_kernel', 'grid_build_kernel', 'pairs_from_grid_kernel',
                         'pair_intersect', 'rows_from_pairs_kernel', 'match_counts_kernel')
        step = tuple(k for k in agg if k.startswith(step_prefixes))
        tot = sum(sum(v) for k, v in agg.items() if k in step)
        md = ['# ncu launch list, round 2 (%s): crop layout, 1,000 C2 images per launch' % tag, '',
              '`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 python bench.py --steps 2 --warmup 3 '
              '--no-graph --no-e2e --no-cpu --no-span --no-c5 --no-check`', '',
              'Per-launch times under ncu are cold-cache and serialised: compare SHARES with `roofline.kernel_share` of '
              'the bench line (paint = flat decode, which also clears the dense matrices, + list kernel; rows = grid build + the '
              'three join kernels; counts = match_counts).', '',
              '| kernel | launches | mean us | total us | share of the step kernels |', '|---|---|---|---|---|']
        for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
            md.append('| `%s` | %d | %.1f | %.1f | %s |' % (k, len(v), sum(v) / len(v), sum(v),
                                                          ('%.3f' % (sum(v) / tot)) if k in step else 'setup / other'))
        open(os.path.join(PROF, 'launches_r02_crop.md'), 'w').write('\n'.join(md) + '\n')

    for name in ('default', 'reference', 'c1', 'c3', 'c4', 'c4_40'):
        src = os.path.join(OUT, 'bench_%s_%s.json' % (tag, name))
        if os.path.exists(src) and os.path.getsize(src):
            shutil.copy(src, os.path.join(PROF, 'bench_r02_%s.json' % name))


if __name__ == '__main__':
    main()
