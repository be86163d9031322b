#!/bin/bash
out=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > $out/t_r02k.log; tail -3 $out/t_r02k.log
for mc in 384 256 128 64; do
AMPIS_ROWS_GRID_MIN_COLS=$mc python bench.py --config c1_powder_example --steps 8 --no-c5 --no-cpu --no-span --no-e2e --no-check > $out/c1_r02k_$mc.json 2> $out/c1_r02k_$mc.err
done
python bench.py --steps 8 --no-c5 --no-cpu --no-span --no-check > $out/c2_r02k.json 2> $out/c2_r02k.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/c?_r02k*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        e = d.get('e2e')
        print('%-24s resident %.3f ms (paint %.3f rows %.3f) %s' % (f.split('/')[-1], d['ms_per_step'], ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step'], ('e2e %.3f ms api %.0f img/s' % (e['ms_per_step'], d['e2e_api']['images_per_s'])) if e else ''))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
