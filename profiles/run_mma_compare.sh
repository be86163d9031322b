#!/bin/bash
# dense-contraction (tcgen05) vs culled AND+popc on the crowded config and on C2; run under gpurun
timeout 200 python -m pytest tests -m gpu -x -q -k tensor_core 2>&1 | tail -3
for K in rows mma; do
  timeout 300 python bench.py --config dense_overlap --images 200 --kernel $K --layout span --no-cpu --no-span --no-e2e --steps 5 > gpurun_out/dense_${K}_span.log 2>&1
done
timeout 300 python bench.py --images 37 --kernel mma --layout full --no-cpu --no-span --no-e2e --steps 3 > gpurun_out/c2_mma_full.log 2>&1
python profiles/show.py gpurun_out/dense_*_span.log gpurun_out/c2_mma_full.log
