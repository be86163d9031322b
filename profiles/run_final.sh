#!/bin/bash
# Round-end evidence on one B200: default bench line (+ reference arm), the other configs, ncu launch lists and
# full-set captures of the hot kernels of every layout, C4 sparse captures, function-level table.
#   gpurun --timeout 1800 -- bash profiles/run_final.sh r01k
TAG=${1:-r01k}
OUT=gpurun_out
python bench.py > $OUT/bench_${TAG}_default.json 2> $OUT/bench_${TAG}_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
bash profiles/run_configs.sh > $OUT/configs_${TAG}.txt 2>&1
python profiles/bench_functions.py > $OUT/functions_${TAG}.jsonl 2> $OUT/functions_${TAG}.err
bash profiles/run_ncu.sh $TAG "full span crop" > $OUT/run_ncu_${TAG}.txt 2>&1
C4="python bench.py --config c4_spheroidite --steps 2 --warmup 3 --images 40 --layout crop --sparse --no-e2e --no-cpu --no-span"
$C4 > $OUT/plain_${TAG}_c4sparse.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_${TAG}_c4sparse.csv $C4 > $OUT/ncu_list_${TAG}_c4sparse.log 2>&1
$C4 > $OUT/plain2_${TAG}_c4sparse.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:intersect_rows_grid -s 3 -c 2 -f -o $OUT/rows_${TAG}_c4sparse $C4 > $OUT/ncu_rows_${TAG}_c4sparse.log 2>&1
$C4 > $OUT/plain3_${TAG}_c4sparse.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:paint -s 3 -c 2 -f -o $OUT/paint_${TAG}_c4sparse $C4 > $OUT/ncu_paint_${TAG}_c4sparse.log 2>&1
python profiles/show.py $OUT/bench_${TAG}_default.json $OUT/cfg_c1.log $OUT/cfg_c3.log $OUT/cfg_c4.log
tail -c 600 $OUT/bench_${TAG}_reference.json
