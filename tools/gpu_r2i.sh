#!/bin/bash
# pass I: generating producers - two alternating groups vs one group (A/B on C3 and C4), full GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r2i_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_tests.log
for grp in 2 1; do
  B2CTR_GEN_GROUPS=$grp timeout 600 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2i_c3_g$grp.json 2> gpurun_out/r2i_c3_g$grp.err
  B2CTR_GEN_GROUPS=$grp timeout 600 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2i_c4_g$grp.json 2> gpurun_out/r2i_c4_g$grp.err
done
tail -3 gpurun_out/r2i_tests.log
