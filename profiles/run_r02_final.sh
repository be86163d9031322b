#!/bin/bash
# what the driver runs at round end, on the final code: GPU tests, smoke(), both bench arms
out=gpurun_out; tag=${1:-final}
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 2 --warmup 0 > $out/bench_${tag}_reference.json 2> $out/bench_${tag}_reference.err
python bench.py --gpus 1 > $out/bench_${tag}_default.json 2> $out/bench_${tag}_default.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_${tag}_default.json').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/bench_${tag}_reference.json').read().strip().splitlines()[-1])
print('ours', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['step']['frac'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
print('reference', r['value'], r.get('impl'), r.get('cpu_baseline', {}).get('cores'))
print('e2e speed-up over the reference arm: %.0f' % (d['e2e']['value'] / r['value']))
PY
