#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "attention" 2>&1 | tail -15
timeout 120 python tools/attn_probe.py 2>&1 | tee gpurun_out/attn_probe_bw64.log
BFK_ATTN_BW=128 timeout 120 python tools/attn_probe.py 2>&1 | tee gpurun_out/attn_probe_bw128.log
