CMD="python bench.py --layout crop --no-cpu --no-span --steps 2 --warmup 3 --images 250"
$CMD > gpurun_out/plain_e2e_crop.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_e2e_crop.csv $CMD > gpurun_out/ncu_list_e2e_crop.log 2>&1
tail -c 400 gpurun_out/plain_e2e_crop.log
