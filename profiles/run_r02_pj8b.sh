#!/bin/bash
out=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python profiles/bench_functions.py --md $out/functions_r02y.md 2>&1 | tail -1 | cut -c1-60; head -12 $out/functions_r02y.md | tail -6 | cut -c1-110
python bench.py --config c3_satellites --images 200 --no-cpu --no-span --no-c5 --no-api > $out/pj8b_c3.json 2> $out/pj8b_c3.err
python bench.py --config c1_powder_example --no-cpu --no-span --no-c5 --no-api > $out/pj8b_c1.json 2> $out/pj8b_c1.err
python - <<PY
import json
for n in ('c3', 'c1'):
    d = json.loads(open('gpurun_out/pj8b_%s.json' % n).read().strip().splitlines()[-1])
    print(n, d['ms_per_step'], d['images_per_s'], d['e2e']['ms_per_step'], d['e2e']['images_per_s'], d['oracle_check']['equal'])
PY
