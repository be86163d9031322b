#!/bin/bash
# zero stores spread over the passes; graph replay as the default of the headline run (all configs)
out=gpurun_out; tag=${1:-q}
timeout 900 python -m pytest tests -x -q -m gpu -k "grid_pruned or batch_pipeline or randomised_batches or native" 2>&1 | tail -2
python bench.py --steps 10 --no-graph --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_plain.json 2> $out/${tag}_r02_c2_plain.err
python bench.py --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c2_graph.json 2> $out/${tag}_r02_c2_graph.err
python bench.py --config c1_powder_example --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c1.json 2> $out/${tag}_r02_c1.err
python bench.py --config c3_satellites --images 200 --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c3.json 2> $out/${tag}_r02_c3.err
python bench.py --config c4_spheroidite --images 160 --sparse --steps 5 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4.json 2> $out/${tag}_r02_c4.err
python bench.py --config c4_spheroidite --images 40 --sparse --steps 10 --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4_40.json 2> $out/${tag}_r02_c4_40.err
python bench.py --config c4_spheroidite --images 40 --sparse --steps 10 --no-graph --no-e2e --no-cpu --no-span --no-c5 > $out/${tag}_r02_c4_40_plain.json 2> $out/${tag}_r02_c4_40_plain.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        print(f.split('/')[-1], d['ms_per_step'], 'paint %.3f rows %.3f' % (ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step']), (d.get('oracle_check') or {}).get('equal'), d['run'].get('cuda_graph'), d['run'].get('cuda_graph_error'), d['images_per_s'])
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
