#!/bin/bash
# ncu --set full of the string decode kernel on the e2e path (250 images per call); usage: <tag>
TAG=${1:-r02}
OUT=gpurun_out
CMDE="python bench.py --steps 2 --warmup 3 --no-cpu --no-span --no-c5 --no-check --no-api"
$CMDE > $OUT/plain_${TAG}_e2e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'rle_string_decode' -s 6 -c 1 -f -o $OUT/strdec_${TAG} $CMDE > $OUT/ncu_strdec_${TAG}.log 2>&1
tail -2 $OUT/ncu_strdec_${TAG}.log
