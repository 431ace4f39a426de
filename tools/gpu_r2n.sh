#!/bin/bash
# pass N (1 GPU): skinny kernels (vec4 TN, unrolled N, warp-per-output split-K reduce), fit() stages its first batch inline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2n_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2n_c2.json 2> gpurun_out/r2n_c2.err
timeout 600 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2n_c3.json 2> gpurun_out/r2n_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2n_c4.json 2> gpurun_out/r2n_c4.err
B2CTR_GEN_GROUPS=2 timeout 600 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2n_c4_g2.json 2> gpurun_out/r2n_c4_g2.err
for c in c2 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2n_launches_$c.csv \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2n_ncu_$c.log 2>&1
done
tail -3 gpurun_out/r2n_tests.log
