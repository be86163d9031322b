#!/bin/bash
# candidate walk compiled for 10 / 12 / 16 CTAs per SM
out=gpurun_out; tag=${1:-g}
AMPIS_PJ_BLOCKS=16 timeout 600 python -m pytest tests -x -q -m gpu -k "grid_pruned or randomised_batches or native" 2>&1 | tail -2
for b in 10 12 16; do
AMPIS_PJ_BLOCKS=$b python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_b$b.json 2> $out/${tag}_r02_c2_b$b.err
AMPIS_PJ_BLOCKS=$b python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4_b$b.json 2> $out/${tag}_r02_c4_b$b.err
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
