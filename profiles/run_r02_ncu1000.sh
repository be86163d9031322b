#!/bin/bash
# ncu --set full of the crop-layout kernels at the bench's launch size (1,000 C2 images); usage: <tag>
TAG=${1:-r02}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-span --no-c5 --no-check"
$CMD > $OUT/plain_${TAG}_crop1000.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'rle_flat|rle_measure_paint_list|grid_build|pairs_from_grid|pair_intersect|rows_from_pairs|match_counts' \
    -s 21 -c 7 -f -o $OUT/crop1000_${TAG} $CMD > $OUT/ncu_crop1000_${TAG}.log 2>&1
tail -3 $OUT/ncu_crop1000_${TAG}.log
