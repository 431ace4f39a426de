#!/bin/bash
# pass O: knock-out sweep of the persistent tcgen05 GEMM (which part of a short-K tile costs the ~4 us?)
mkdir -p gpurun_out
for d in 0 1 2 3 4 8 12 15; do
  B2CTR_TC_DEBUG=$d timeout 300 python tools/gemm_sweep.py >> gpurun_out/r2o_gemm_sweep.log 2>> gpurun_out/r2o_gemm_sweep.err
done
B2CTR_TC_TMA=0 timeout 300 python tools/gemm_sweep.py >> gpurun_out/r2o_gemm_sweep.log 2>> gpurun_out/r2o_gemm_sweep.err
cat gpurun_out/r2o_gemm_sweep.log
