#!/bin/bash
# pass H: coalesced (row, quad) generator, once-per-lookup id validation, skinny wgrad slices, Feeder steady-state path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2h_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_tests.log
timeout 300 python bench.py > gpurun_out/r2h_bench_c2.json 2> gpurun_out/r2h_bench_c2.err
timeout 300 python bench.py --dist zipf --no-cpu-baseline > gpurun_out/r2h_bench_c2_zipf.json 2> gpurun_out/r2h_bench_c2_zipf.err
timeout 600 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2h_bench_c3.json 2> gpurun_out/r2h_bench_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err
R=/tmp/ncu_reports; mkdir -p $R
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 $NCU --set full --import-source on -k regex:gemm_planes_ws -s 49 -c 1 -f -o $R/cin1 $B --config c3 > gpurun_out/r2h_prof_c3.log 2>&1
ncu -i $R/cin1.ncu-rep --page raw --csv > gpurun_out/r2h_cin1_raw.csv 2>/dev/null
ncu -i $R/cin1.ncu-rep --page source --csv > gpurun_out/r2h_cin1_source.csv 2>/dev/null
tail -3 gpurun_out/r2h_tests.log; du -sh gpurun_out
