#!/bin/bash
export BFK_GEMM_2CTA=1
echo "== gemm tests (2cta on)"; timeout -k 10 120 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" -p no:cacheprovider 2>&1 | tail -15
echo "== diag"; timeout -k 10 100 python tools/diag_gemm.py 2>&1 | tail -34
echo "== epilogue bench (2cta on)"; timeout -k 10 100 python tools/bench_epilogue.py 2>&1 | tail -12
