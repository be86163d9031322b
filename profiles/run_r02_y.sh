#!/bin/bash
# join passes: entry walk with four loads in flight, AND+popc with 2 / 4 columns in flight
out=gpurun_out; tag=${1:-y}
timeout 1500 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised or native or crop or many_images or one_call or golden_matching or satellites or sparse" 2>&1 | tail -3
for c in 2 4; do
AMPIS_PI_COLS=$c python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_pi$c.json 2> $out/${tag}_r02_c2_pi$c.err
AMPIS_PI_COLS=$c python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4_pi$c.json 2> $out/${tag}_r02_c4_pi$c.err
AMPIS_PI_COLS=$c python bench.py --config c3_satellites --images 200 --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c3_pi$c.json 2> $out/${tag}_r02_c3_pi$c.err
done
AMPIS_E2E_WORKERS=6 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w6.json 2> $out/${tag}_r02_e2e_w6.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        ks = d['roofline']['kernel_share']
        print(f.split('/')[-1], d['ms_per_step'], 'paint %.3f rows %.3f' % (ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step']), e.get('ms_per_step'), (d.get('oracle_check') or {}).get('equal'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-span --no-c5 --no-check"
$CMD > $out/plain_${tag}_crop1000.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/launches_${tag}_crop1000.csv $CMD > $out/ncu_list_${tag}.log 2>&1
tail -1 $out/ncu_list_${tag}.log | cut -c1-100
