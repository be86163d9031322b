#!/bin/bash
# string decode rewrite: parity tests, then the e2e legs
out=gpurun_out; tag=${1:-u}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --no-cpu --no-span --no-c5 --no-check > $out/${tag}_r02_e2e.json 2> $out/${tag}_r02_e2e.err
AMPIS_E2E_WORKERS=6 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w6.json 2> $out/${tag}_r02_e2e_w6.err
AMPIS_E2E_WORKERS=8 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w8.json 2> $out/${tag}_r02_e2e_w8.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d['e2e']
        print(f.split('/')[-1], d['ms_per_step'], e['ms_per_step'], e['images_per_s'], json.dumps(d.get('e2e_api'))[:300])
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
