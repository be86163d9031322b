#!/bin/bash
# e2e legs of C1 / C3 / C4 with calls of ~250,000 masks; calls in flight auto and fixed
out=gpurun_out; tag=${1:-p}
for w in 0 2 4; do
AMPIS_E2E_WORKERS=$w python bench.py --config c4_spheroidite --images 160 --sparse --no-c5 --no-span --no-cpu --no-check --no-api > $out/${tag}_r02_c4_w$w.json 2> $out/${tag}_r02_c4_w$w.err
AMPIS_E2E_WORKERS=$w python bench.py --config c3_satellites --images 200 --no-c5 --no-span --no-cpu --no-check --no-api > $out/${tag}_r02_c3_w$w.json 2> $out/${tag}_r02_c3_w$w.err
done
python bench.py --config c1_powder_example --no-c5 --no-span --no-cpu --no-check --no-api > $out/${tag}_r02_c1_w0.json 2> $out/${tag}_r02_c1_w0.err
python bench.py --no-c5 --no-span --no-cpu --no-check --no-api > $out/${tag}_r02_c2_w0.json 2> $out/${tag}_r02_c2_w0.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f.split('/')[-1], d['ms_per_step'], d['roofline']['frac'], e.get('ms_per_step'), e.get('images_per_s'), e.get('calls_per_step'), e.get('images_per_call'), e.get('calls_in_flight'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
