#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 120 python tools/attn_one.py > gpurun_out/attn_one.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_.*v2 -s 4 -c 4 -o gpurun_out/attn_v2 python tools/attn_one.py > gpurun_out/attn_ncu.log 2>&1
echo "rc=$?" >> gpurun_out/attn_ncu.log; tail -n 5 gpurun_out/attn_ncu.log
