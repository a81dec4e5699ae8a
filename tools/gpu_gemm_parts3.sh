#!/bin/bash
for d in 0 16 4 8; do BFK_GEMM_DBG=$d timeout 120 python tools/bench_gemm_parts.py 2>&1 | grep bfk | tee -a gpurun_out/gemm_parts3.log; done
