#!/bin/bash
# round 2 A/B on the culled (crop-layout) step: fused decode flat (rle_flat.cu) vs lane groups (rle_paint.cu, r1),
# grid-pruned rows as the three-pass join (intersect_pairs.cu) vs the single rows kernel (intersect_grid.cu, r1)
# usage: bash profiles/run_r02_ab.sh <tag> ["group/grid flat/grid flat/pairs"]
tag=${1:-r02a}
combos=${2:-"group/grid flat/grid flat/pairs"}
for c in $combos; do
  dec=${c%/*}; rk=${c#*/}
  AMPIS_CROP_DECODE=$dec AMPIS_ROWS_KERNEL=$rk bash profiles/run_grid_quick.sh ${tag}_${dec}_${rk} 2>&1 | tail -3
done
