#!/bin/bash
mkdir -p gpurun_out
echo "== gemm tests (cluster on)"; timeout -k 10 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm" -p no:cacheprovider 2>&1 | tail -5
echo "== diag (cluster on)"; timeout -k 10 200 python tools/diag_gemm.py 2>&1 | tail -12
echo "== epilogue bench (cluster on)"; timeout -k 10 200 python tools/bench_epilogue.py 2>&1 | tail -12
echo "== epilogue bench (cluster off)"; BFK_GEMM_CLUSTER=0 timeout -k 10 200 python tools/bench_epilogue.py 2>&1 | tail -12
