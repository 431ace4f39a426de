#!/bin/bash
# pass J (1 GPU): full GPU suite; C3 with the rolling-prefetch CIN generator; C4 with the zero-row skip; C5's per-GPU
# footprint (26 x 12.5M x 128 = 166 GB of tables) on ONE GPU before spending 8 GPUs on it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r2j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_tests.log
timeout 600 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2j_c3.json 2> gpurun_out/r2j_c3.err
timeout 600 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2j_c4.json 2> gpurun_out/r2j_c4.err
timeout 900 python bench.py --config c5 --no-cpu-baseline --steps 10 > gpurun_out/r2j_c5_n1.json 2> gpurun_out/r2j_c5_n1.err
nvidia-smi --query-gpu=memory.total,memory.used --format=csv > gpurun_out/r2j_mem.txt
tail -3 gpurun_out/r2j_tests.log; tail -c 400 gpurun_out/r2j_c5_n1.err
