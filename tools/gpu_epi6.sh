#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "specialised" 2>&1 | tail -3
timeout 120 python tools/bench_epilogue.py 2>&1 | tee gpurun_out/bench_epilogue6.log
