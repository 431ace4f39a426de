#!/bin/bash
# DRAM bytes of the fused gather / scatter on the last tree (the source of profiles/traffic.json)
mkdir -p gpurun_out
B2CTR_STEP_GRAPH=0 timeout 300 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  -k regex:uniform_ -s 8 -c 2 --csv --log-file gpurun_out/r2w_embed_traffic.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_ncu.log 2>&1
tail -8 gpurun_out/r2w_embed_traffic.csv
