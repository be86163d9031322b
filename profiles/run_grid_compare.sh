#!/bin/bash
# Crop rows kernel: scan of all column boxes vs grid-pruned candidates (+ sparse output), per config.
#   gpurun -- bash profiles/run_grid_compare.sh r01h
tag=${1:-r01h}
out=gpurun_out
for cfg_n in "c4_spheroidite 40" "c3_satellites 200" "c2_powder_batch 1000"; do
  set -- $cfg_n
  for k in scan grid; do
    python bench.py --config $1 --images $2 --layout crop --kernel $k --no-cpu --no-span --steps 10 --warmup 3 \
      > $out/grid_${tag}_$1_$k.json 2> $out/grid_${tag}_$1_$k.err
  done
  python bench.py --config $1 --images $2 --layout crop --kernel grid --sparse --no-cpu --no-span --steps 10 --warmup 3 \
      > $out/grid_${tag}_$1_sparse.json 2> $out/grid_${tag}_$1_sparse.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/grid_*_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'FAILED', e); continue
    print('%-50s img/s %9.0f ms/step %8.3f share %s e2e %s' % (f.split('/')[-1], d['images_per_s'], d['ms_per_step'],
          {k: round(v, 3) for k, v in d['roofline']['kernel_share'].items()}, d.get('e2e', {}).get('value')))
PY
