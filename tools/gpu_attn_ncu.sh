#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/attn_probe.py 2>&1 | tee gpurun_out/attn_probe.log
ATTN_SHAPE=fusion timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 12 -c 4 -o gpurun_out/attn_fusion python tools/attn_probe.py > gpurun_out/ncu_attn.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn.log
