#!/bin/bash
# crowded frames (dense_overlap), sorted tiles as the drop-in API runs them: single CTA vs CTA pair; ncu of the pair kernel
out=gpurun_out
for k in mma mma2; do
  timeout 300 python bench.py --config dense_overlap --images 512 --kernel $k --mma-sort --layout span --no-cpu --no-span --no-e2e --steps 10 --warmup 3 > $out/pair_dos_$k.json 2> $out/pair_dos_$k.err
done
python profiles/show.py $out/pair_dos_mma.json $out/pair_dos_mma2.json
ncu --set full --clock-control none --import-source on -k regex:intersect_mma_pair -s 3 -c 2 -f -o $out/mma2_r01j python bench.py --images 37 --kernel mma2 --layout full --no-cpu --no-span --no-e2e --steps 2 --warmup 3 > $out/ncu_mma2.log 2>&1
