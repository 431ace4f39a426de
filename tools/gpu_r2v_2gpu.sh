#!/bin/bash
# 2 GPUs, last tree: sharded parity tests + the default bench line under torchrun (what the driver's scaling run does)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sharded_gpu.py -q > gpurun_out/r2v_sharded_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_sharded_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2v_c2_n2.json 2> gpurun_out/r2v_c2_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2v_ref_n2.json 2> gpurun_out/r2v_ref_n2.err
tail -2 gpurun_out/r2v_sharded_tests.log; head -c 200 gpurun_out/r2v_c2_n2.json; echo; head -c 200 gpurun_out/r2v_ref_n2.json
