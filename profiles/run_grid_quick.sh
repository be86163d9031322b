#!/bin/bash
# quick A/B of the grid kernel: C2 dense, C4 sparse, C3 dense (crop layout)
tag=${1:-x}
out=gpurun_out
python bench.py --config c2_powder_batch --images 1000 --layout crop --kernel grid --no-cpu --no-span --no-e2e > $out/gq_${tag}_c2.json 2> $out/gq_${tag}_c2.err
python bench.py --config c4_spheroidite --images 40 --layout crop --kernel grid --sparse --no-cpu --no-span --no-e2e > $out/gq_${tag}_c4s.json 2> $out/gq_${tag}_c4s.err
python bench.py --config c3_satellites --images 200 --layout crop --kernel grid --no-cpu --no-span --no-e2e > $out/gq_${tag}_c3.json 2> $out/gq_${tag}_c3.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/gq_${tag}_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    ks = d['roofline']['kernel_share']
    print('%-30s img/s %9.0f ms/step %8.3f rows_ms %.3f paint_ms %.3f' % (f.split('/')[-1], d['images_per_s'], d['ms_per_step'], ks['rows'] * d['ms_per_step'], ks['paint'] * d['ms_per_step']))
PY
