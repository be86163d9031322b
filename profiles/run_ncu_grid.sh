#!/bin/bash
# ncu --set full of the grid-pruned crop rows kernel (C2 frames and C4 sparse) and of the crop decode kernel.
#   gpurun -- bash profiles/run_ncu_grid.sh r01h
set -u
TAG=${1:-r01h}
OUT=gpurun_out
C2="python bench.py --steps 2 --warmup 3 --images 182 --sub 91 --layout crop --kernel grid --no-e2e --no-cpu --no-span"
C4="python bench.py --config c4_spheroidite --steps 2 --warmup 3 --images 40 --layout crop --sparse --no-e2e --no-cpu --no-span"
$C2 > $OUT/plain_${TAG}_gridc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:intersect_rows_grid -s 6 -c 2 \
    -f -o $OUT/rows_${TAG}_gridc2 $C2 > $OUT/ncu_rows_${TAG}_gridc2.log 2>&1
$C4 > $OUT/plain_${TAG}_gridc4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:intersect_rows_grid -s 3 -c 2 \
    -f -o $OUT/rows_${TAG}_gridc4 $C4 > $OUT/ncu_rows_${TAG}_gridc4.log 2>&1
$C4 > $OUT/plain2_${TAG}_gridc4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file $OUT/launches_${TAG}_gridc4.csv $C4 > $OUT/ncu_list_${TAG}_gridc4.log 2>&1
$C2 > $OUT/plain2_${TAG}_gridc2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:paint -s 6 -c 2 \
    -f -o $OUT/paint_${TAG}_crop $C2 > $OUT/ncu_paint_${TAG}_crop.log 2>&1
ls $OUT | grep $TAG
