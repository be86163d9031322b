#!/bin/bash
# two GPUs: NCCL tests + the bench under torchrun (what the driver's SCALE step launches)
out=gpurun_out
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu 2>&1 | tail -5 > $out/t_r02_n2.log; tail -3 $out/t_r02_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_r02_n2.json 2> $out/bench_r02_n2.err
tail -c 600 $out/bench_r02_n2.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_r02_n2.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'img/s %.0f' % d['images_per_s'], 'ms %.3f' % d['ms_per_step'], 'e2e %.0f img/s' % d['e2e']['images_per_s'], 'api %.0f' % d['e2e_api']['images_per_s'], 'c5 %.3f ms %.0f img/s' % (d['c5_strong']['ms_per_pass'], d['c5_strong']['images_per_s']), 'full %.0f' % d['full_layout']['images_per_s'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $out/bench_r02_n2_reference.json 2> $out/bench_r02_n2_reference.err
head -c 300 $out/bench_r02_n2_reference.json
