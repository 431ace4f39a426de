#!/bin/bash
# pass B: A/B of the TMA producers and the persisting-L2 window on C2 (device-resident value only) + debug of the C3 capture
mkdir -p gpurun_out
for tma in 1 0; do for per in 0 1; do
  B2CTR_TC_TMA=$tma B2CTR_L2_PERSIST=$per python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2b_c2_tma${tma}_persist${per}.json 2> gpurun_out/r2b_c2_tma${tma}_persist${per}.err
done; done
B2CTR_L2_PERSIST=0 python tools/gpu_dbg.py > gpurun_out/r2b_dbg.log 2>&1
B2CTR_L2_PERSIST=0 python -m pytest tests/test_baseline_shapes_gpu.py -q > gpurun_out/r2b_shapes.log 2>&1
tail -3 gpurun_out/r2b_shapes.log
