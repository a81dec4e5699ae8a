#!/bin/bash
mkdir -p gpurun_out
ATTN_SHAPE=vision timeout 60 python tools/attn_probe.py > gpurun_out/attn_plain.log 2>&1 &&
ATTN_SHAPE=vision timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 4 -o gpurun_out/attn_final_vision python tools/attn_probe.py > gpurun_out/ncu_attn_v.log 2>&1
echo "ncu exit $?"
ATTN_SHAPE=fusion timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 9 -c 4 -o gpurun_out/attn_final_fusion python tools/attn_probe.py > gpurun_out/ncu_attn_f.log 2>&1
echo "ncu exit $?"
