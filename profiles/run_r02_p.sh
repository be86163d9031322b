#!/bin/bash
# 8 GPUs, e2e leg only: spinning vs blocking wait, 4 vs 2 calls in flight per rank
out=gpurun_out
for v in "1 4" "0 2" "1 2"; do
  set -- $v
  AMPIS_E2E_BLOCKING=$1 AMPIS_E2E_WORKERS=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29540+$1*10+$2)) bench.py --gpus 8 --steps 10 --no-cpu --no-span --no-c5 --no-api --no-check > $out/p_r02_b$1_w$2.json 2> $out/p_r02_b$1_w$2.err
  python -c "
import json
d=json.loads(open('gpurun_out/p_r02_b$1_w$2.json').read().strip().splitlines()[-1])
print('blocking=$1 workers=$2: resident %.3f ms; e2e %.3f ms %.0f img/s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['images_per_s']))
"
done
