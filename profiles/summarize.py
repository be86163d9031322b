#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

    python profiles/summarize.py <tag>      # e.g. r01

Reads  gpurun_out/launches_<tag>_{full,span,crop}.csv         (ncu --metrics gpu__time_duration.sum)
       gpurun_out/{paint,rows}_<tag>_{full,span,crop}.ncu-rep (ncu --set full)
       gpurun_out/mma_<tag>.ncu-rep                           (ncu --set full of the tcgen05 contraction)
Writes profiles/launches_<tag>_<layout>.md, profiles/kernels_<tag>.md and updates profiles/traffic.json
(per-launch dram bytes of the hot kernels, read by bench.py for roofline.traffic).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

UNITS = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3,
         'msecond': 1e-3, 'second': 1, 'nsecond': 1e-9}

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
           'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
           'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
           'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def read_csv_after_header(path):
    lines = open(path).read().splitlines()
    for i, l in enumerate(lines):
        if l.startswith('"ID"'):
            return list(csv.DictReader(io.StringIO('\n'.join(lines[i:]))))
    return []


def launches(tag, layout):
    path = os.path.join(OUT, 'launches_%s_%s.csv' % (tag, layout))
    if not os.path.exists(path):
        return None
    rows = read_csv_after_header(path)
    agg = collections.OrderedDict()
    for r in rows:
        k = r['Kernel Name'].split('(')[0]
        agg.setdefault(k, []).append(float(r['Metric Value']))
    tot = sum(sum(v) for v in agg.values())
    md = ['# ncu launch list, %s, layout %s' % (tag, layout), '',
          '`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 '
          '--images 182 --sub 91 --layout %s --no-e2e --no-cpu --no-span`' % layout, '',
          'Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py\'s '
          '`roofline.kernel_share`, not absolutes.', '',
          '| kernel | launches | mean us | total us | share |', '|---|---|---|---|---|']
    for k, v in agg.items():
        md.append('| `%s` | %d | %.1f | %.1f | %.3f |' % (k[:70], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3,
                                                        sum(v) / tot))
    md.append('')
    md.append('total device time %.1f us over %d launches' % (tot / 1e3, len(rows)))
    open(os.path.join(PROF, 'launches_%s_%s.md' % (tag, layout)), 'w').write('\n'.join(md) + '\n')
    return agg


def rep_metrics(path):
    r = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {'kernel': row[hdr.index('Kernel Name')].split('(')[0]}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                try:
                    d[m] = float(row[i].replace(',', '')) * UNITS.get(units[i], 1)
                except ValueError:
                    d[m] = row[i]
        out.append(d)
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    tpath = os.path.join(PROF, 'traffic.json')
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    md = ['# ncu --set full summaries, %s' % tag, '',
          'Command: `ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 6 -c 2 python '
          'bench.py --steps 2 --warmup 3 --images 182 --sub 91 --layout <layout> --no-e2e --no-cpu --no-span` '
          '(c2_powder_batch: 91 images = 91,000 masks of 1024x1024 per launch).', '']
    for layout in ('full', 'span', 'crop'):
        launches(tag, layout)
        for kern in ('paint', 'rows'):
            path = os.path.join(OUT, '%s_%s_%s.ncu-rep' % (kern, tag, layout))
            if not os.path.exists(path):
                continue
            ms = rep_metrics(path)
            if not ms:
                continue
            md += ['## %s, layout %s' % (ms[0]['kernel'], layout), '', '| metric | ' + ' | '.join(
                'launch %d' % i for i in range(len(ms))) + ' |', '|---|' + '---|' * len(ms)]
            for m in METRICS:
                if m in ms[0]:
                    md.append('| %s | ' % m + ' | '.join(
                        ('%.6g' % x[m]) if isinstance(x.get(m), float) else str(x.get(m)) for x in ms) + ' |')
            rd = sum(x['dram__bytes_read.sum'] for x in ms) / len(ms)
            wr = sum(x['dram__bytes_write.sum'] for x in ms) / len(ms)
            t = sum(x['gpu__time_duration.sum'] for x in ms) / len(ms)
            md += ['', 'per launch: DRAM read %.1f MB + write %.1f MB = %.1f MB in %.3f ms -> %.0f GB/s' % (
                rd / 1e6, wr / 1e6, (rd + wr) / 1e6, t * 1e3, (rd + wr) / t / 1e9), '']
            traffic['c2_powder_batch/%s/%s' % (layout, kern)] = {
                'bytes_per_launch': rd + wr, 'images_per_launch': 91, 'bytes_per_image': (rd + wr) / 91,
                'source': 'profiles/kernels_%s.md' % tag}
    mpath = os.path.join(OUT, 'mma_%s.ncu-rep' % tag)
    if os.path.exists(mpath):
        extra = ['sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active',
                 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
                 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
        METRICS.extend(m for m in extra if m not in METRICS)
        ms = rep_metrics(mpath)
        if ms:
            md += ['## %s (int8 tcgen05 contraction), C2 frames' % ms[0]['kernel'], '',
                   'Command (profiles/run_ncu_mma.sh): `ncu --set full --clock-control none --import-source on -k '
                   'regex:intersect_mma -s 3 -c 2 python bench.py --images 37 --kernel mma --layout full --no-cpu '
                   '--no-span --no-e2e --steps 2 --warmup 3` (37 images x 8 tiles of 128x256 = 296 CTAs = two waves, '
                   '8192 slabs of 128 pixels per tile).', '',
                   '| metric | ' + ' | '.join('launch %d' % i for i in range(len(ms))) + ' |', '|---|' + '---|' * len(ms)]
            for m in METRICS:
                if m in ms[0]:
                    md.append('| %s | ' % m + ' | '.join(
                        ('%.6g' % x[m]) if isinstance(x.get(m), float) else str(x.get(m)) for x in ms) + ' |')
            t = sum(x['gpu__time_duration.sum'] for x in ms) / len(ms)
            ops = 2.0 * 37 * 250000 * 1024 * 1024
            md += ['', 'per launch: %.3f ms for 2*G*P*H*W = %.3g integer ops -> %.0f TOP/s algorithmic (%.0f executed with '
                   'the 512x512 tile padding); tensor pipe (IMMA) active %.1f %% of cycles; DRAM %.0f MB.' % (
                       t * 1e3, ops, ops / t / 1e12, ops / t / 1e12 * (512 * 512) / (500 * 500),
                       ms[0]['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'],
                       (ms[0]['dram__bytes_read.sum'] + ms[0]['dram__bytes_write.sum']) / 1e6),
                   'SASS (`cuobjdump -sass ampis_b200/libampis_b200.so`): `UTCIMMA` (tcgen05.mma kind::i8), `UTCBAR` '
                   '(tcgen05.commit), `LDTM.x32` (tcgen05.ld), `UBLKCP` (cp.async.bulk in the rows kernels), '
                   '`SYNCS.*` (mbarrier).', '']
    open(os.path.join(PROF, 'kernels_%s.md' % tag), 'w').write('\n'.join(md) + '\n')
    json.dump(traffic, open(tpath, 'w'), indent=1, sort_keys=True)
    print('\n'.join(md))


if __name__ == '__main__':
    main()
