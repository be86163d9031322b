#!/bin/bash
# round 2, call g: tests, crossover incl. the TMA tiled kernel, e2e with pinned endpoints, launch list of a whole bench run
out=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > $out/t_r02g.log; tail -3 $out/t_r02g.log
python profiles/crossover.py > $out/crossover_r02.jsonl 2> $out/crossover_r02.err; tail -3 $out/crossover_r02.err; cat $out/crossover_r02.jsonl | cut -c1-420
for w in 1 2 3; do
  for chunk in 250 500; do
    AMPIS_E2E_WORKERS=$w python bench.py --steps 5 --warmup 3 --e2e-chunk $chunk --no-cpu --no-span --no-c5 --no-api --no-check \
        > $out/e2e_r02g_w${w}_c${chunk}.json 2> $out/e2e_r02g_w${w}_c${chunk}.err
  done
done
python - <<PY
import glob, json
for f in sorted(glob.glob('gpurun_out/e2e_r02g_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d['e2e']
        print('%-34s resident %.3f ms  e2e %.3f ms (wall %.3f)  %8.0f img/s  ratio %.2f' % (f.split('/')[-1], d['ms_per_step'], e['ms_per_step'], e['wall_ms_per_step'], e['images_per_s'], d['ms_per_step'] / e['ms_per_step']))
    except Exception as ex:
        print(f, 'FAILED', ex)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-span --no-c5 --no-api --no-check --e2e-chunk 500"
$CMD > $out/plain_r02g.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_r02g.csv $CMD > $out/ncu_list_r02g.log 2>&1
tail -2 $out/ncu_list_r02g.log | cut -c1-200
