#!/bin/bash
# single-CTA vs CTA-pair tensor-core contraction on C2 frames (37 images per launch = two waves) and on the crowded config
out=gpurun_out
for k in mma mma2; do
  timeout 300 python bench.py --images 37 --kernel $k --layout span --no-cpu --no-span --no-e2e --steps 10 --warmup 3 > $out/pair_c2_$k.json 2> $out/pair_c2_$k.err
  timeout 300 python bench.py --images 37 --kernel $k --mma-sort --layout span --no-cpu --no-span --no-e2e --steps 10 --warmup 3 > $out/pair_c2s_$k.json 2> $out/pair_c2s_$k.err
  timeout 300 python bench.py --config dense_overlap --images 512 --kernel $k --layout span --no-cpu --no-span --no-e2e --steps 10 --warmup 3 > $out/pair_do_$k.json 2> $out/pair_do_$k.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/pair_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e, open(f.replace('.json', '.err')).read()[-400:]); continue
    r = d['roofline']
    print('%-28s img/s %9.0f ms/step %8.3f rows_ms/launch %.3f  %s %.1f %s frac %.3f' % (f.split('/')[-1], d['images_per_s'], d['ms_per_step'], r['launch_ms'], r['bound'], r['achieved'], r['unit'], r['frac']))
PY
