out=gpurun_out
for k in scan grid; do
python bench.py --config c1_powder_example --images 2000 --layout crop --kernel $k --no-cpu --no-span --no-e2e > $out/c1_$k.json 2> $out/c1_$k.err
done
python profiles/show.py $out/c1_scan.json $out/c1_grid.json
AMPIS_ROWS_GRID_MIN_COLS=1 python profiles/profile_call.py 2>&1 | head -3
