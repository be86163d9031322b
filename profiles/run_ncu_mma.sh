#!/bin/bash
# ncu capture of the tcgen05 contraction on C2-sized images (37 images = 296 tiles = 2 waves of 148 CTAs)
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --images 37 --kernel mma --layout full --no-cpu --no-span --no-e2e --steps 2 --warmup 3"
$CMD > $OUT/plain_mma_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:intersect_mma -s 3 -c 2 \
    -f -o $OUT/mma_$TAG $CMD > $OUT/ncu_mma_$TAG.log 2>&1
ls $OUT | grep mma_$TAG
