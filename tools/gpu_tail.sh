#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "heatmap or argmax or bce" -p no:cacheprovider 2>&1 | tail -4
timeout -k 10 600 python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout -k 10 300 python tools/bench_tail.py > gpurun_out/tail.log 2>&1; head -12 gpurun_out/tail.log
