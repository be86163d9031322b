#!/bin/bash
# the bench lines of the round without the ncu captures; usage: <tag>
TAG=${1:-r02}
OUT=gpurun_out
python bench.py > $OUT/bench_${TAG}_default.json 2> $OUT/bench_${TAG}_default.err
python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
python bench.py --config c1_powder_example --no-c5 > $OUT/bench_${TAG}_c1.json 2> $OUT/bench_${TAG}_c1.err
python bench.py --config c3_satellites --images 200 --no-c5 > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err
python bench.py --config c4_spheroidite --images 160 --sparse --no-c5 --no-span > $OUT/bench_${TAG}_c4.json 2> $OUT/bench_${TAG}_c4.err
python bench.py --config c4_spheroidite --images 40 --sparse --no-c5 --no-span --no-cpu --no-e2e > $OUT/bench_${TAG}_c4_40.json 2> $OUT/bench_${TAG}_c4_40.err
python bench.py --no-graph --no-cpu --no-span --no-c5 --no-e2e > $OUT/bench_${TAG}_nograph.json 2> $OUT/bench_${TAG}_nograph.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ls -la $OUT | grep bench_${TAG} | tail -20
