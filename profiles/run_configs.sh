#!/bin/bash
# the other BASELINE.json configs through bench.py (C1 powder example, C3 satellites, C4 spheroidite)
timeout 600 python bench.py --config c1_powder_example --images 2000 --steps 5 --no-cpu > gpurun_out/cfg_c1.log 2> gpurun_out/cfg_c1.err
timeout 600 python bench.py --config c3_satellites --images 100 --steps 5 --no-cpu > gpurun_out/cfg_c3.log 2> gpurun_out/cfg_c3.err
timeout 600 python bench.py --config c4_spheroidite --images 40 --steps 5 --sparse --no-cpu > gpurun_out/cfg_c4.log 2> gpurun_out/cfg_c4.err
tail -n 3 gpurun_out/cfg_*.err
python profiles/show.py gpurun_out/cfg_c1.log gpurun_out/cfg_c3.log gpurun_out/cfg_c4.log
