#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "attention or layernorm or colsum or gemm" 2>&1 | tail -8
timeout 120 python tools/attn_probe.py 2>&1 | tee gpurun_out/attn_probe3.log
BFK_ATTN_FWD_BW=128 timeout 120 python tools/attn_probe.py 2>&1 | grep fwd | tee gpurun_out/attn_probe3_fwd128.log
