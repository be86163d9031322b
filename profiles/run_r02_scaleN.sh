#!/bin/bash
# one point of the 1/2/4/8-GPU sweep (the box is charged per GPU: each N on a box of its own size); usage: <N> [more N...]
out=gpurun_out
for n in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --no-cpu --no-span > $out/scale_r02v_n$n.json 2> $out/scale_r02v_n$n.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_r02v_n$n.json').read().strip().splitlines()[-1])
    print('N=%d resident %.0f img/s %.3f ms | e2e %.0f img/s %.3f ms | api %.0f img/s | c5 %.3f ms' % (
        d['n_gpus'], d['images_per_s'], d['ms_per_step'], d['e2e']['images_per_s'], d['e2e']['ms_per_step'],
        d['e2e_api']['images_per_s'], d['c5_strong']['ms_per_pass']))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/scale_r02v_n$n.err').read()[-800:])
PY
done
