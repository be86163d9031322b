#!/bin/bash
out=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > $out/t_r02l.log; tail -3 $out/t_r02l.log
python profiles/bench_functions.py --md $out/functions_r02.md > $out/functions_r02.jsonl 2> $out/functions_r02.err; tail -3 $out/functions_r02.err; head -30 $out/functions_r02.md | cut -c1-200
python bench.py > $out/bench_r02l_default.json 2> $out/bench_r02l_default.err
python bench.py --config c1_powder_example --no-c5 > $out/bench_r02l_c1.json 2> $out/bench_r02l_c1.err
python -c "
import json
for n in ('default','c1'):
    d=json.loads(open('gpurun_out/bench_r02l_%s.json'%n).read().strip().splitlines()[-1])
    print(n, '%.0f img/s %.3f ms; e2e %.0f img/s %.3f ms; api %.0f img/s; frac %.3f step %.3f' % (d['images_per_s'], d['ms_per_step'], d['e2e']['images_per_s'], d['e2e']['ms_per_step'], d['e2e_api']['images_per_s'], d['roofline']['frac'], d['roofline']['step']['frac']))
"
