#!/bin/bash
# pass E: validate the CIN fold epilogue + flattened sequence scatter, bench C3 / C4, then ncu evidence
# (B200_PROFILING.md recipe: launch lists of one step per config + --set full captures of the dominant kernels;
#  numbers printed by bench.py under ncu are never bench values)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2e_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
timeout 600 python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err
NCU="ncu --clock-control none"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
for c in c2 c3 c4; do
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 1200 --csv --log-file gpurun_out/r2_launches_$c.csv $B --config $c > gpurun_out/r2_prof_$c.log 2>&1
done
timeout 600 $NCU --set full --import-source on -k regex:gather_uniform_fwd -s 4 -c 1 -o gpurun_out/r2_full_gather $B --config c2 >> gpurun_out/r2_prof_c2.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:scatter_uniform_bwd -s 4 -c 1 -o gpurun_out/r2_full_scatter $B --config c2 >> gpurun_out/r2_prof_c2.log 2>&1
timeout 600 $NCU --set full -k regex:gemm_planes_ws -s 36 -c 9 -o gpurun_out/r2_full_gemm $B --config c2 >> gpurun_out/r2_prof_c2.log 2>&1
timeout 900 $NCU --set full -k regex:gemm_planes_ws -s 40 -c 10 -o gpurun_out/r2_full_cin $B --config c3 >> gpurun_out/r2_prof_c3.log 2>&1
timeout 900 $NCU --set full -k regex:gemm_planes_ws -s 30 -c 8 -o gpurun_out/r2_full_att $B --config c4 >> gpurun_out/r2_prof_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2e_tests.log
