#!/bin/bash
# row pass of the join with four pairs' loads in flight
out=gpurun_out; tag=${1:-j3}
timeout 900 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised or native or many_images or one_call or golden or satellites or sparse" 2>&1 | tail -2
python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2.json 2> $out/${tag}_r02_c2.err
python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4.json 2> $out/${tag}_r02_c4.err
python bench.py --config c3_satellites --images 200 --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c3.json 2> $out/${tag}_r02_c3.err
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
