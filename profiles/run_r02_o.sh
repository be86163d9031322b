#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "many_images or c_caller or one_call or batch_equals or native" 2>&1 | tail -4 > $out/t_r02o.log; tail -2 $out/t_r02o.log
for i in 1 2; do
python bench.py --steps 10 --no-cpu --no-span --no-c5 --no-check > $out/o_r02_$i.json 2> $out/o_r02_$i.err
python -c "
import json
d=json.loads(open('gpurun_out/o_r02_$i.json').read().strip().splitlines()[-1])
print('resident %.3f ms; e2e %.3f ms %.0f img/s; api %.0f img/s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['images_per_s'], d['e2e_api']['images_per_s']))
"
done
