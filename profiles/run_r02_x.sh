#!/bin/bash
# string decode variants: parity tests, e2e legs, launch list of the e2e path
out=gpurun_out; tag=${1:-x}
timeout 1500 python -m pytest tests -x -q -m gpu -k "string or fixtures or golden or many_images or one_call or c_caller or det_seg or get_ddicts or real" 2>&1 | tail -3
AMPIS_E2E_WORKERS=6 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w6.json 2> $out/${tag}_r02_e2e_w6.err
AMPIS_E2E_WORKERS=4 python bench.py --no-cpu --no-span --no-c5 --no-check --no-api > $out/${tag}_r02_e2e_w4.json 2> $out/${tag}_r02_e2e_w4.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/${tag}_r02_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        print(f.split('/')[-1], d['ms_per_step'], e.get('ms_per_step'), e.get('images_per_s'))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-400:])
PY
CMDE="python bench.py --steps 2 --warmup 3 --no-cpu --no-span --no-c5 --no-check --no-api"
$CMDE > $out/plain_${tag}_e2e.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_${tag}_e2e.csv $CMDE > $out/ncu_list_${tag}_e2e.log 2>&1
tail -1 $out/ncu_list_${tag}_e2e.log | cut -c1-200
