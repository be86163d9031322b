#!/bin/bash
# dense matrices zeroed by the decode kernel (AMPIS_ZERO_WITH_DECODE) vs beside the join
out=gpurun_out; tag=${1:-zz}
timeout 900 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised or native or crop or golden_matching or full_size" 2>&1 | tail -2
for z in 0 1; do
AMPIS_ZERO_WITH_DECODE=$z python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_z$z.json 2> $out/${tag}_r02_c2_z$z.err
AMPIS_ZERO_WITH_DECODE=$z python bench.py --steps 10 --graph --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2g_z$z.json 2> $out/${tag}_r02_c2g_z$z.err
AMPIS_ZERO_WITH_DECODE=$z python bench.py --config c1_powder_example --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c1_z$z.json 2> $out/${tag}_r02_c1_z$z.err
AMPIS_ZERO_WITH_DECODE=$z python bench.py --config c3_satellites --images 200 --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c3_z$z.json 2> $out/${tag}_r02_c3_z$z.err
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
