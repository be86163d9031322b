#!/bin/bash
# AND+popc pass of the join: eight lanes per pair (AMPIS_PI_FLAT=0) vs lanes over the concatenated columns of eight pairs (1)
out=gpurun_out; tag=${1:-f}
AMPIS_PI_FLAT=1 timeout 1200 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised or native or crop or many_images or one_call or golden_matching or satellites or sparse or full_size" 2>&1 | tail -2
AMPIS_PI_FLAT=2 timeout 1200 python -m pytest tests -x -q -m gpu -k "grid_pruned or randomised_batches or native or crop_decode" 2>&1 | tail -2
for f in 0 1 2; do
AMPIS_PI_FLAT=$f python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_f$f.json 2> $out/${tag}_r02_c2_f$f.err
AMPIS_PI_FLAT=$f python bench.py --config c1_powder_example --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c1_f$f.json 2> $out/${tag}_r02_c1_f$f.err
AMPIS_PI_FLAT=$f python bench.py --config c3_satellites --images 200 --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c3_f$f.json 2> $out/${tag}_r02_c3_f$f.err
AMPIS_PI_FLAT=$f python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4_f$f.json 2> $out/${tag}_r02_c4_f$f.err
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
