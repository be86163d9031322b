#!/bin/bash
# round 2 evidence: default bench line, the other configs, reference arm, ncu launch list + full-set capture of the
# crop-layout kernels at the bench's launch size and at 91 images per launch (comparable with round 1)
# usage: bash profiles/run_r02_evidence.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
python bench.py > $OUT/bench_${TAG}_default.json 2> $OUT/bench_${TAG}_default.err
python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
python bench.py --config c1_powder_example --no-c5 > $OUT/bench_${TAG}_c1.json 2> $OUT/bench_${TAG}_c1.err
python bench.py --config c3_satellites --images 200 --no-c5 > $OUT/bench_${TAG}_c3.json 2> $OUT/bench_${TAG}_c3.err
python bench.py --config c4_spheroidite --images 160 --sparse --no-c5 --no-span > $OUT/bench_${TAG}_c4.json 2> $OUT/bench_${TAG}_c4.err
python bench.py --config c4_spheroidite --images 40 --sparse --no-c5 --no-span --no-cpu --no-e2e > $OUT/bench_${TAG}_c4_40.json 2> $OUT/bench_${TAG}_c4_40.err
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu --no-span --no-c5 --no-check"
$CMD > $OUT/plain_${TAG}_crop1000.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_${TAG}_crop1000.csv $CMD > $OUT/ncu_list_${TAG}.log 2>&1
$CMD > $OUT/plain2_${TAG}_crop1000.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'rle_flat|rle_measure_paint_list|grid_build|pairs_from_grid|pair_intersect|rows_from_pairs|match_counts' \
    -s 21 -c 7 -f -o $OUT/crop1000_${TAG} $CMD > $OUT/ncu_crop1000_${TAG}.log 2>&1
CMD91="python bench.py --steps 2 --warmup 3 --no-graph --images 182 --sub 91 --no-e2e --no-cpu --no-span --no-c5 --no-check"
$CMD91 > $OUT/plain_${TAG}_crop91.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'rle_flat|rle_measure_paint_list|grid_build|pairs_from_grid|pair_intersect|rows_from_pairs|match_counts' \
    -s 42 -c 7 -f -o $OUT/crop91_${TAG} $CMD91 > $OUT/ncu_crop91_${TAG}.log 2>&1
CMDE="python bench.py --steps 2 --warmup 3 --no-cpu --no-span --no-c5 --no-check --no-api"
$CMDE > $OUT/plain_${TAG}_e2e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rle_string_decode' -s 6 -c 1 -f -o $OUT/strdec_${TAG} $CMDE > $OUT/ncu_strdec_${TAG}.log 2>&1
ls -la $OUT | grep ${TAG} | tail -30
