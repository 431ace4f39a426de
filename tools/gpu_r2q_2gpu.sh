#!/bin/bash
# 2 GPUs, C5's per-GPU footprint (26 x 12.5M x 128 per GPU): NVLink peer loads / red.add against the NCCL all-to-all
# transport.  At 8 GPUs the peer kernels ran 40x slower per byte than on one GPU's own 166 GB (r2g_c5_n8.json).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --config c5 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/r2q_c5_n2_peer.json 2> gpurun_out/r2q_c5_n2_peer.err
B2CTR_SHARD_MODE=a2a timeout 600 $TR --master-port 29532 bench.py --gpus 2 --config c5 --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/r2q_c5_n2_a2a.json 2> gpurun_out/r2q_c5_n2_a2a.err
tail -c 300 gpurun_out/r2q_c5_n2_peer.err; tail -c 300 gpurun_out/r2q_c5_n2_a2a.err
