#!/bin/bash
# pass D: re-validate the generated-operand GEMMs (8 producer warps), BN tail A/B, bench C2 / Zipf / C3 / C4
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
timeout 300 python bench.py > gpurun_out/r2d_bench_c2.json 2> gpurun_out/r2d_bench_c2.err
B2CTR_TC_BN_TAIL=0 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2d_c2_notail.json 2> gpurun_out/r2d_c2_notail.err
timeout 300 python bench.py --dist zipf --no-cpu-baseline > gpurun_out/r2d_bench_c2_zipf.json 2> gpurun_out/r2d_bench_c2_zipf.err
timeout 600 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err
tail -4 gpurun_out/r2d_tests.log
