#!/bin/bash
# round-2 first GPU pass: full GPU test-suite, bench lines of C2 / C3 / C4 (+ Zipf), L2 persistence limits
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
python - > gpurun_out/r2a_l2.txt 2>&1 <<'PY'
from cuda import cudart
for name in ("cudaDevAttrL2CacheSize", "cudaDevAttrMaxPersistingL2CacheSize", "cudaDevAttrMaxAccessPolicyWindowSize",
             "cudaDevAttrMultiProcessorCount", "cudaDevAttrMaxSharedMemoryPerBlockOptin"):
    print(name, cudart.cudaDeviceGetAttribute(getattr(cudart.cudaDeviceAttr, name), 0))
PY
python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
python bench.py > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
python bench.py --config c2 --dist zipf --no-cpu-baseline > gpurun_out/r2a_bench_c2_zipf.json 2> gpurun_out/r2a_bench_c2_zipf.err
python bench.py --config c3 --no-cpu-baseline > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
python bench.py --config c4 --no-cpu-baseline > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err
tail -5 gpurun_out/r2a_tests.log
