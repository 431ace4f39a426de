#!/bin/bash
# 2 GPUs: sharded parity tests on the final tree (automatic transport choice in set_dist)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -q > gpurun_out/r2t_sharded_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_sharded_tests.log
tail -3 gpurun_out/r2t_sharded_tests.log
