#!/bin/bash
timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "gemm" 2>&1 | tail -4
timeout 120 python tools/bench_gemm_parts.py 2>&1 | grep -v cuBLAS | tee gpurun_out/gemm_parts9.log
