#!/bin/bash
out=gpurun_out
for z in 0 1 2; do
AMPIS_ZERO_AHEAD=$z python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-span --no-c5 --no-check > $out/z_r02i_$z.json 2> $out/z_r02i_$z.err
AMPIS_ZERO_AHEAD=$z python bench.py --steps 8 --warmup 3 --graph --no-e2e --no-cpu --no-span --no-c5 --no-check > $out/z_r02i_${z}g.json 2> $out/z_r02i_${z}g.err
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/z_r02i_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = d['roofline']['kernel_share']
        print('%-30s resident %.3f ms (paint %.3f rows %.3f)' % (f.split('/')[-1], d['ms_per_step'], ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step']))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
