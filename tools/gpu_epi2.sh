#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "gemm" -p no:cacheprovider 2>&1 | tail -3
timeout -k 10 300 python tools/bench_epilogue.py > gpurun_out/epi.log 2>&1; cat gpurun_out/epi.log
