#!/bin/bash
# e2e legs after the decode change: default workers / chunk, and a few variants
out=gpurun_out
python bench.py --no-cpu --no-span --no-c5 --no-check > $out/t_r02_e2e.json 2> $out/t_r02_e2e.err
AMPIS_E2E_WORKERS=6 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/t_r02_e2e_w6.json 2> $out/t_r02_e2e_w6.err
python bench.py --no-cpu --no-span --no-c5 --no-check --no-api --e2e-chunk 125 > $out/t_r02_e2e_c125.json 2> $out/t_r02_e2e_c125.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/t_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], d['ms_per_step'], json.dumps(d['e2e'])[:600])
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
