#!/bin/bash
# Profiling recipe (run under gpurun, one GPU).  Usage: bash profiles/run_ncu.sh <tag> [layouts]
# 1) plain run must exit 0, 2) launch list (per-launch device time), 3) --set full on the two hot kernels.
set -u
TAG=${1:-r01}
LAYOUTS=${2:-"full span"}
OUT=gpurun_out
for LAYOUT in $LAYOUTS; do
  CMD="python bench.py --steps 2 --warmup 3 --images 182 --sub 91 --layout $LAYOUT --no-e2e --no-cpu --no-span"
  $CMD > $OUT/plain_${TAG}_$LAYOUT.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/launches_${TAG}_$LAYOUT.csv $CMD > $OUT/ncu_list_${TAG}_$LAYOUT.log 2>&1
  $CMD > $OUT/plain2_${TAG}_$LAYOUT.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:paint -s 6 -c 2 \
      -f -o $OUT/paint_${TAG}_$LAYOUT $CMD > $OUT/ncu_paint_${TAG}_$LAYOUT.log 2>&1
  $CMD > $OUT/plain3_${TAG}_$LAYOUT.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:intersect_rows -s 6 -c 2 \
      -f -o $OUT/rows_${TAG}_$LAYOUT $CMD > $OUT/ncu_rows_${TAG}_$LAYOUT.log 2>&1
done
ls $OUT | grep $TAG
