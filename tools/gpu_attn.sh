#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
BFK_ATTN_V2_FWD=1 ATTN_STAGE=fwd timeout 150 python tools/attn_check.py > gpurun_out/attn_fwd.log 2>&1; echo "rc=$?" >> gpurun_out/attn_fwd.log
ATTN_STAGE=bwd timeout 150 python tools/attn_check.py > gpurun_out/attn_bwd.log 2>&1; echo "rc=$?" >> gpurun_out/attn_bwd.log
BFK_ATTN_V2=0 ATTN_STAGE=all timeout 150 python tools/attn_check.py > gpurun_out/attn_old.log 2>&1; echo "rc=$?" >> gpurun_out/attn_old.log
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/smi_after.log 2>&1
tail -n 25 gpurun_out/attn_fwd.log gpurun_out/attn_bwd.log; tail -n 12 gpurun_out/attn_old.log
