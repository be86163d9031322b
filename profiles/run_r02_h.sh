#!/bin/bash
# round 2, call h: zero-ahead + K=5 decode, pipelined e2e
out=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > $out/t_r02h.log; tail -3 $out/t_r02h.log
for w in 2 3 4; do
  for chunk in 125 250 500; do
    AMPIS_E2E_WORKERS=$w python bench.py --steps 8 --warmup 3 --e2e-chunk $chunk --no-cpu --no-span --no-c5 --no-api --no-check \
        > $out/e2e_r02h_w${w}_c${chunk}.json 2> $out/e2e_r02h_w${w}_c${chunk}.err
  done
done
AMPIS_ZERO_AHEAD=0 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-span --no-c5 --no-check > $out/e2e_r02h_nozero.json 2> $out/e2e_r02h_nozero.err
python bench.py --steps 8 --warmup 3 --graph --no-e2e --no-cpu --no-span --no-c5 --no-check > $out/e2e_r02h_graph.json 2> $out/e2e_r02h_graph.err
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/e2e_r02h_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e')
        ks = d['roofline']['kernel_share']
        msg = '%-30s resident %.3f ms (paint %.3f rows %.3f)' % (f.split('/')[-1], d['ms_per_step'], ks['paint'] * d['ms_per_step'], ks['rows'] * d['ms_per_step'])
        if e:
            msg += '  e2e %.3f ms (wall %.3f)  %8.0f img/s  ratio %.2f' % (e['ms_per_step'], e['wall_ms_per_step'], e['images_per_s'], d['ms_per_step'] / e['ms_per_step'])
        print(msg)
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
tail -3 $out/e2e_r02h_w3_c250.err
