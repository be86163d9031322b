#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised_batches or native or crop_layout" 2>&1 | tail -3
for z in 0 1; do
AMPIS_ZERO_BESIDE_JOIN=$z python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/r_r02_z$z.json 2> $out/r_r02_z$z.err
AMPIS_ZERO_BESIDE_JOIN=$z python bench.py --steps 10 --graph --no-e2e --no-cpu --no-span --no-c5 > $out/r_r02_z${z}g.json 2> $out/r_r02_z${z}g.err
AMPIS_ZERO_BESIDE_JOIN=$z python bench.py --config c1_powder_example --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/r_r02_c1_z$z.json 2> $out/r_r02_c1_z$z.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/r_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        print('%-26s resident %.3f ms (paint %.3f rows %.3f) check %s' % (f.split('/')[-1], d['ms_per_step'], ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step'], (d.get('oracle_check') or {}).get('equal')))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
