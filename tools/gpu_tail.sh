#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_policy_gpu.py -x -q -k "heatmap or head_project or bce or argmax or golden_tiny or mask_pick or generic_heatmap" > gpurun_out/t_tail.log 2>&1; tail -n 25 gpurun_out/t_tail.log
timeout 200 python tools/bench_tail.py > gpurun_out/tail_bench.log 2>&1
head -n 12 gpurun_out/tail_bench.log
