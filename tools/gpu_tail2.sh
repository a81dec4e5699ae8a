#!/bin/bash
for v in 0 1 2 3 4 5; do echo "== variant $v"; BFK_HM_VARIANT=$v timeout -k 10 200 python tools/bench_tail.py 2>&1 | grep "heatmap_fwd"; done
