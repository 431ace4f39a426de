#!/bin/bash
# final sanity of the round's last tree: smoke(), full GPU suite, default bench line
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2u_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2u_smoke.log
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2u_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2u_c2.json 2> gpurun_out/r2u_c2.err
tail -2 gpurun_out/r2u_smoke.log; tail -2 gpurun_out/r2u_tests.log; head -c 300 gpurun_out/r2u_c2.json
timeout 300 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2u_c3.json 2> gpurun_out/r2u_c3.err
timeout 300 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2u_c4.json 2> gpurun_out/r2u_c4.err
