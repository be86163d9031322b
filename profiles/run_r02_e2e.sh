#!/bin/bash
# e2e (C ABI from host buffers) vs images per call and calls in flight; crop layout, C2
tag=${1:-r02f}
out=gpurun_out
for w in 1 2 3; do
  for chunk in 125 250 500 1000; do
    AMPIS_E2E_WORKERS=$w python bench.py --steps 5 --warmup 3 --e2e-chunk $chunk --no-cpu --no-span --no-c5 --no-api --no-check \
        > $out/e2e_${tag}_w${w}_c${chunk}.json 2> $out/e2e_${tag}_w${w}_c${chunk}.err
  done
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/e2e_${tag}_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d['e2e']
        print('%-34s resident %.3f ms  e2e %.3f ms (wall %.3f)  %8.0f img/s  ratio %.2f' % (f.split('/')[-1], d['ms_per_step'], e['ms_per_step'], e['wall_ms_per_step'], e['images_per_s'], d['ms_per_step'] / e['ms_per_step']))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
