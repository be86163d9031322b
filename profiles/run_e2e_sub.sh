out=gpurun_out
for sub in 250 125 63; do
  python bench.py --layout crop --sub $sub --no-cpu --no-span > $out/e2e_sub_$sub.json 2> $out/e2e_sub_$sub.err
done
python - <<'PY'
import json
for sub in (250, 125, 63):
    d = json.loads(open('gpurun_out/e2e_sub_%d.json' % sub).read().strip().splitlines()[-1])
    print(sub, 'device ms/step %.3f' % d['ms_per_step'], 'e2e ms/step %.3f' % d['e2e']['ms_per_step'], 'e2e G pairs/s %.1f' % (d['e2e']['value'] / 1e9))
PY
