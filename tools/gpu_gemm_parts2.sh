#!/bin/bash
export BFK_GEMM_2CTA=0
for d in 0 2 6 10 14; do BFK_GEMM_DBG=$d timeout 120 python tools/bench_gemm_parts2.py 2>&1 | tee -a gpurun_out/gemm_parts2.log; done
