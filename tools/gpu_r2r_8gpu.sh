#!/bin/bash
# 8 GPUs: C5 again with the automatic transport choice (1.16 TB of peer-mapped shards -> NCCL all-to-all)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 8 --config c5 --steps 10 --no-cpu-baseline > gpurun_out/r2r_c5_n8_auto.json 2> gpurun_out/r2r_c5_n8_auto.err
tail -c 500 gpurun_out/r2r_c5_n8_auto.err
