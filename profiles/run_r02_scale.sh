#!/bin/bash
# 1/2/4/8-GPU sweep of the default bench (what the driver's SCALE step runs), on an 8-GPU box
out=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $out/t_r02_scale.log; tail -2 $out/t_r02_scale.log
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu > $out/scale_r02_n$n.json 2> $out/scale_r02_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --no-cpu > $out/scale_r02_n$n.json 2> $out/scale_r02_n$n.err
  fi
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open('gpurun_out/scale_r02_n%d.json' % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, 'FAILED', e); continue
    if base is None: base = d
    print('N=%d  resident %.0f img/s (%.2fx)  %.3f ms | e2e %.0f img/s (%.2fx) %.3f ms | api %.0f img/s | c5 %.3f ms (%.2fx) | full %.0f img/s' % (
        n, d['images_per_s'], d['images_per_s'] / base['images_per_s'], d['ms_per_step'], d['e2e']['images_per_s'],
        d['e2e']['images_per_s'] / base['e2e']['images_per_s'], d['e2e']['ms_per_step'], d['e2e_api']['images_per_s'],
        d['c5_strong']['ms_per_pass'], base['c5_strong']['ms_per_pass'] / d['c5_strong']['ms_per_pass'], d['full_layout']['images_per_s']))
PY
nvidia-smi topo -m > $out/topo_r02.txt 2>&1; head -12 $out/topo_r02.txt
