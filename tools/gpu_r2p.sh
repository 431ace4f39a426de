#!/bin/bash
# pass P: CTA-pair TMA loads signal the leader's mbarrier directly (no relay warp): GEMM tests first (own timeout: a wrong
# barrier protocol hangs), then the sweep, then the step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_abi_dense_gpu.py tests/test_layers_gpu.py -q -x > gpurun_out/r2p_gemm_tests.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/r2p_gemm_tests.log
tail -3 gpurun_out/r2p_gemm_tests.log
if [ $rc -ne 0 ]; then exit 1; fi
for d in 0 15; do B2CTR_TC_DEBUG=$d timeout 200 python tools/gemm_sweep.py >> gpurun_out/r2p_gemm_sweep.log 2>&1; done
cat gpurun_out/r2p_gemm_sweep.log
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/r2p_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_tests.log; tail -2 gpurun_out/r2p_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2p_c2.json 2> gpurun_out/r2p_c2.err
timeout 300 python bench.py --config c3 --no-cpu-baseline --no-e2e > gpurun_out/r2p_c3.json 2> gpurun_out/r2p_c3.err
timeout 300 python bench.py --config c4 --no-cpu-baseline --no-e2e > gpurun_out/r2p_c4.json 2> gpurun_out/r2p_c4.err
