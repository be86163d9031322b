#!/bin/bash
# flat decode kernel variants: parity tests that exercise it, then the culled step on C2 / C1 / C4 sparse
out=gpurun_out; tag=${1:-s}
timeout 1200 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised or native or crop or many_images or one_call or golden_matching" 2>&1 | tail -3
python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2.json 2> $out/${tag}_r02_c2.err
python bench.py --config c1_powder_example --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c1.json 2> $out/${tag}_r02_c1.err
python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4.json 2> $out/${tag}_r02_c4.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        print('%-22s %.3f ms (paint %.3f rows %.3f) check %s %s' % (f.split('/')[-1], d['ms_per_step'], ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step'], (d.get('oracle_check') or {}).get('equal'), json.dumps(d.get('c4_sparse'))[:200]))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
