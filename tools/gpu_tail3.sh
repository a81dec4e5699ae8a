#!/bin/bash
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "heatmap or argmax or bce" -p no:cacheprovider 2>&1 | tail -3
timeout -k 10 300 python -m pytest tests/test_policy_gpu.py -m gpu -q -k "tiny or inference or generic" -p no:cacheprovider 2>&1 | tail -3
for v in 0 1; do echo "== variant $v"; BFK_HM_VARIANT=$v timeout -k 10 200 python tools/bench_tail.py 2>&1 | grep "heatmap\|head_proj" | grep -v "^{"; done
