#!/bin/bash
# candidate walk of the join: a thread per row (AMPIS_PJ_LANES=1) vs eight lanes per row (8)
out=gpurun_out; tag=${1:-pj8}
AMPIS_PJ_LANES=8 timeout 900 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised or native or many_images or one_call or golden_matching or satellites or sparse" 2>&1 | tail -2
for l in 1 8; do
AMPIS_PJ_LANES=$l python bench.py --steps 10 --no-cpu --no-span --no-c5 --no-api > $out/${tag}_r02_c2_l$l.json 2> $out/${tag}_r02_c2_l$l.err
AMPIS_PJ_LANES=$l python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4_l$l.json 2> $out/${tag}_r02_c4_l$l.err
AMPIS_PJ_LANES=$l python bench.py --config c3_satellites --images 200 --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c3_l$l.json 2> $out/${tag}_r02_c3_l$l.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        e = d.get('e2e') or {}
        print(f.split('/')[-1], d['ms_per_step'], 'paint %.3f rows %.3f' % (ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step']), (d.get('oracle_check') or {}).get('equal'), e.get('ms_per_step'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
