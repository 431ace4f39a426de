#!/bin/bash
# 2 GPUs: multi-GPU parity tests + C2 / C5-shaped (small) bench lines on the row-sharded path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -q > gpurun_out/r2f_sharded_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_sharded_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2f_c2_n2.json 2> gpurun_out/r2f_c2_n2.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --config c5small --no-cpu-baseline > gpurun_out/r2f_c5small_n2.json 2> gpurun_out/r2f_c5small_n2.err
tail -3 gpurun_out/r2f_sharded_tests.log
