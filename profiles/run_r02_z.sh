#!/bin/bash
# AND+popc pass variants (AMPIS_PI_VARIANT): registers / CTAs per SM and descriptor prefetch
out=gpurun_out; tag=${1:-z}
timeout 600 python -m pytest tests -x -q -m gpu -k "grid_pruned or randomised_batches" 2>&1 | tail -2
for v in 0 1 2 3 4 5; do
AMPIS_PI_VARIANT=$v python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_v$v.json 2> $out/${tag}_r02_c2_v$v.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        print(f.split('/')[-1], d['ms_per_step'], 'paint %.3f rows %.3f' % (ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step']), (d.get('oracle_check') or {}).get('equal'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
