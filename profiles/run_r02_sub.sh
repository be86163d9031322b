#!/bin/bash
# crop-layout step vs images per launch (C2 dense, C4 sparse), current default kernels
tag=${1:-r02c}
out=gpurun_out
for sub in 91 250 1000; do
  python bench.py --config c2_powder_batch --images 1000 --sub $sub --layout crop --no-cpu --no-span --no-e2e > $out/sub_${tag}_c2_$sub.json 2> $out/sub_${tag}_c2_$sub.err
done
python bench.py --config c2_powder_batch --images 1000 --sub 1000 --layout crop --graph --no-cpu --no-span --no-e2e > $out/sub_${tag}_c2_1000g.json 2> $out/sub_${tag}_c2_1000g.err
python bench.py --config c4_spheroidite --images 40 --layout crop --sparse --no-cpu --no-span --no-e2e > $out/sub_${tag}_c4s_40.json 2> $out/sub_${tag}_c4s_40.err
python bench.py --config c4_spheroidite --images 160 --sub 160 --layout crop --sparse --no-cpu --no-span --no-e2e > $out/sub_${tag}_c4s_160.json 2> $out/sub_${tag}_c4s_160.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/sub_${tag}_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    ks = d['roofline']['kernel_share']
    print('%-30s img/s %9.0f ms/step %8.3f rows_ms %.3f paint_ms %.3f counts_ms %.3f crop_frac %.3f' % (f.split('/')[-1], d['images_per_s'], d['ms_per_step'], ks['rows'] * d['ms_per_step'], ks['paint'] * d['ms_per_step'], ks['counts'] * d['ms_per_step'], d['cropped_accounting']['frac']))
PY
