#!/usr/bin/env python
"""Print the key numbers of bench.py JSON lines: python profiles/show.py file..."""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'unreadable', e)
        continue
    for name, k in (('main', d), ('span', d.get('span_layout')), ('crop', d.get('crop_layout'))):
        if not k:
            continue
        r = k['roofline']
        print('%-28s %-5s img/s %8.0f  ms/step %7.3f  %s frac %.3f launch_ms %.4f  canon %.2f  %s  e2e %s' % (
            f.split('/')[-1], d['config']['layout'] if name == 'main' else name, k['images_per_s'], k['ms_per_step'],
            r['kernel'].replace('_kernel', ''), r['frac'], r['launch_ms'], r['step_canonical']['frac'],
            {a: round(b, 3) for a, b in r['kernel_share'].items()},
            ('%.2fG' % (k['e2e']['value'] / 1e9)) if k.get('e2e') else '-'))
