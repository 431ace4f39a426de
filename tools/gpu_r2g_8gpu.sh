#!/bin/bash
# 8 GPUs: BASELINE.json configs[4] (C5: 26 x 100M-row tables row-sharded, emb_dim 128, batch 262144) and C2 at N = 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2g_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --config c5 --no-cpu-baseline > gpurun_out/r2g_c5_n8.json 2> gpurun_out/r2g_c5_n8.err
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2g_c2_n8.json 2> gpurun_out/r2g_c2_n8.err
tail -c 600 gpurun_out/r2g_c5_n8.err
