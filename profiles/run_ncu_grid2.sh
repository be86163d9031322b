set -u
OUT=gpurun_out
C2="python bench.py --steps 2 --warmup 3 --images 182 --sub 91 --layout crop --kernel grid --no-e2e --no-cpu --no-span"
C4="python bench.py --config c4_spheroidite --steps 2 --warmup 3 --images 40 --layout crop --sparse --no-e2e --no-cpu --no-span"
ncu --set full --clock-control none --import-source on -k regex:intersect_rows_grid -s 6 -c 1 -f -o $OUT/rows_r01i_gridc2 $C2 > $OUT/ncu_rows_r01i_gridc2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:intersect_rows_grid -s 3 -c 1 -f -o $OUT/rows_r01i_gridc4 $C4 > $OUT/ncu_rows_r01i_gridc4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches_r01i_gridc4.csv $C4 > $OUT/ncu_list_r01i_gridc4.log 2>&1
