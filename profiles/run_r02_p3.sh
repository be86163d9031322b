#!/bin/bash
# e2e: fewer, bigger calls
out=gpurun_out; tag=${1:-p3}
run() { name=$1; shift; w=$1; shift; AMPIS_E2E_WORKERS=$w python bench.py "$@" --no-c5 --no-span --no-cpu --no-check --no-api > $out/${tag}_r02_$name.json 2> $out/${tag}_r02_$name.err; }
run c2_1000_w2 2 --e2e-chunk 1000
run c2_500_w3 3 --e2e-chunk 500
run c2_500_w2 2 --e2e-chunk 500
run c2_334_w4 4 --e2e-chunk 334
run c4_80_w3 3 --config c4_spheroidite --images 160 --sparse --e2e-chunk 80
run c4_80_w2 2 --config c4_spheroidite --images 160 --sparse --e2e-chunk 80
run c4_160_w2 2 --config c4_spheroidite --images 160 --sparse --e2e-chunk 160
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f.split('/')[-1], d['ms_per_step'], e.get('ms_per_step'), e.get('images_per_s'), e.get('calls_per_step'), e.get('images_per_call'), e.get('calls_in_flight'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
