#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2; do BFK_GEMM_DBG=$d timeout 120 python tools/bench_gemm_parts.py 2>&1 | tee -a gpurun_out/gemm_parts.log; done
BFK_GEMM_2CTA=1 timeout 120 python tools/bench_gemm_parts.py 2>&1 | tee -a gpurun_out/gemm_parts.log
