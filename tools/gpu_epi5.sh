#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "gemm" 2>&1 | tail -5
timeout 120 python tools/bench_gemm_parts.py 2>&1 | tee gpurun_out/gemm_parts5.log
timeout 120 python tools/bench_epilogue.py 2>&1 | tee gpurun_out/bench_epilogue5.log
