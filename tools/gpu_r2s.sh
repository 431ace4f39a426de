#!/bin/bash
# final sanity of the round's last tree: smoke(), full GPU suite, default bench line
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2s_smoke.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2s_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_tests.log
timeout 600 python bench.py > gpurun_out/r2s_c2.json 2> gpurun_out/r2s_c2.err
tail -2 gpurun_out/r2s_smoke.log; tail -2 gpurun_out/r2s_tests.log; head -c 300 gpurun_out/r2s_c2.json
