#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; : > gpurun_out/attn_abl.log
for d in 0 1 2 4 8 16 32 3 7 15 63; do BFK_ATTN_DBG=$d timeout 60 python tools/attn_time.py >> gpurun_out/attn_abl.log 2>&1; done
cat gpurun_out/attn_abl.log
