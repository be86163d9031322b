#!/bin/bash
# ncu --set full of the crop-layout step's kernels (one launch of each), after a plain run of the same command
# usage: bash profiles/run_r02_ncu.sh <tag> [extra bench args]
TAG=${1:-r02}
shift
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --images 182 --sub 91 --layout crop --no-e2e --no-cpu --no-span $*"
$CMD > $OUT/plain_${TAG}_crop.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'rle_flat|rle_measure_paint|grid_build|pairs_from_grid|pair_intersect|rows_from_pairs|intersect_rows_grid|match_counts' \
    -s 48 -c 8 -f -o $OUT/crop_${TAG} $CMD > $OUT/ncu_crop_${TAG}.log 2>&1
tail -3 $OUT/ncu_crop_${TAG}.log
