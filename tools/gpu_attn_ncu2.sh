#!/bin/bash
mkdir -p gpurun_out
ATTN_SHAPE=vision timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/attn_fwd_v python tools/attn_probe.py > gpurun_out/ncu_attn_f.log 2>&1
echo "ncu exit $?"
ATTN_SHAPE=vision timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 6 -c 2 -o gpurun_out/attn_bwd_v python tools/attn_probe.py > gpurun_out/ncu_attn_b.log 2>&1
echo "ncu exit $?"
