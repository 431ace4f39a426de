#!/bin/bash
# pass K (1 GPU): capture-invalidation probe, input-pipeline probe, flattened / register-resident generators on C3 + C4,
# full suite, launch lists
mkdir -p gpurun_out
timeout 300 python tools/gpu_dbg_capture.py bf16x3 > gpurun_out/r2k_capture_dbg.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2k_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_tests.log
timeout 300 python tools/feed_probe.py > gpurun_out/r2k_feed_probe.json 2> gpurun_out/r2k_feed_probe.err
timeout 600 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_c3.json 2> gpurun_out/r2k_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err
B2CTR_GEN_GROUPS=2 timeout 600 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2k_c4_g2.json 2> gpurun_out/r2k_c4_g2.err
for c in c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2k_launches_$c.csv \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2k_ncu_$c.log 2>&1
done
tail -3 gpurun_out/r2k_tests.log; tail -5 gpurun_out/r2k_capture_dbg.log
