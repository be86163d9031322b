#!/bin/bash
# persistent string decode (register variants), flat decode with in-warp slices, match_counts: tests, benches, e2e launch list
out=gpurun_out; tag=${1:-w}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2.json 2> $out/${tag}_r02_c2.err
for b in 4 5; do
AMPIS_SD_BLOCKS=$b AMPIS_E2E_WORKERS=6 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w6_b$b.json 2> $out/${tag}_r02_e2e_w6_b$b.err
AMPIS_SD_BLOCKS=$b AMPIS_E2E_WORKERS=4 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w4_b$b.json 2> $out/${tag}_r02_e2e_w4_b$b.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        ks = d['roofline']['kernel_share']
        print(f.split('/')[-1], d['ms_per_step'], 'paint %.3f rows %.3f counts %.3f' % (ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step'], ks['counts'] * d['ms_per_step']), e.get('ms_per_step'), e.get('images_per_s'), (d.get('oracle_check') or {}).get('equal'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
CMDE="python bench.py --steps 2 --warmup 3 --no-cpu --no-span --no-c5 --no-check --no-api"
$CMDE > $out/plain_${tag}_e2e.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_${tag}_e2e.csv $CMDE > $out/ncu_list_${tag}_e2e.log 2>&1
tail -1 $out/ncu_list_${tag}_e2e.log | cut -c1-200
