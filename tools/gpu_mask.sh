#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py -x -q -k "mask_branch or mask_losses or golden_tiny or mask_pick or generic_heatmap" > gpurun_out/t_mask.log 2>&1; tail -n 30 gpurun_out/t_mask.log
