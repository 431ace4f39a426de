#!/bin/bash
# pass L (1 GPU): CIN forward with resident X (j-block-major k order) on C3; fit() timeline; ncu --set full (with source)
# of the two worst small-K GEMMs of a C4 step: the attention-input generator GEMM and the [B*T,80]x[80,40] layer
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2l_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_tests.log
timeout 600 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2l_c3.json 2> gpurun_out/r2l_c3.err
timeout 300 python tools/feed_probe.py > gpurun_out/r2l_feed_probe.json 2> gpurun_out/r2l_feed_probe.err
B2CTR_STEP_GRAPH=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_planes_ws_kernel -s 45 -c 2 \
    -o gpurun_out/r2l_c4_smallk -f python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2l_ncu_c4.log 2>&1
ls -la gpurun_out/r2l_c4_smallk.ncu-rep
tail -3 gpurun_out/r2l_tests.log
