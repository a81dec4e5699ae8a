#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_policy_gpu.py -x -q -k "heatmap or head_project or bce or argmax or golden_tiny or mask_pick or generic_heatmap" > gpurun_out/t_tail.log 2>&1; tail -n 5 gpurun_out/t_tail.log
timeout 200 python tools/bench_tail.py > gpurun_out/tail_bench.log 2>&1
head -n 5 gpurun_out/tail_bench.log
BFK_HM_MINB=3 timeout 200 python tools/bench_tail.py > gpurun_out/tail_bench_minb3.log 2>&1
head -n 2 gpurun_out/tail_bench_minb3.log
timeout 120 python tools/tail_one.py > gpurun_out/tail_one.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:heatmap_fwd_kernel|heatmap_bwd_stage1|head_project_fwd" -s 3 -c 6 -o gpurun_out/tail_r02b python tools/tail_one.py > gpurun_out/tail_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/tail_ncu.log; tail -n 3 gpurun_out/tail_ncu.log
