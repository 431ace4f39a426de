#!/bin/bash
# pass C: GPU test-suite (generated-operand CIN / DIN GEMMs, sorted update, OOB, model goldens), A/B of the TMA
# producers and the persisting-L2 window on C2, bench lines of C3 / C4, capture debug
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x -k "cin_generated or din_first or sorted or out_of_range" > gpurun_out/r2c_new_tests.log 2>&1
echo "pytest(new) rc=$?" >> gpurun_out/r2c_new_tests.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
for tma in 1 0; do for per in 0 1; do
  B2CTR_TC_TMA=$tma B2CTR_L2_PERSIST=$per timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2c_c2_tma${tma}_persist${per}.json 2> gpurun_out/r2c_c2_tma${tma}_persist${per}.err
done; done
timeout 600 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
timeout 300 python tools/gpu_dbg.py > gpurun_out/r2c_dbg.log 2>&1
tail -3 gpurun_out/r2c_new_tests.log; tail -3 gpurun_out/r2c_tests.log
