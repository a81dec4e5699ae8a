#!/bin/bash
# ncu --set full evidence of round 2: GEMM epilogue variants (TMA store / TMA load on and off) and the attention kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in gg mulaux res; do
  timeout 120 python tools/gemm_one.py $v > /dev/null 2>&1 || echo "gemm_one $v failed"
  timeout 400 ncu --set full --clock-control none -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/r02_gemm_${v}_tma python tools/gemm_one.py $v > gpurun_out/ncu_$v.log 2>&1; echo "ncu $v rc=$?"
  BFK_GEMM_TMA_STORE=0 BFK_GEMM_TMA_LOAD=0 timeout 400 ncu --set full --clock-control none -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/r02_gemm_${v}_reg python tools/gemm_one.py $v > gpurun_out/ncu_${v}_reg.log 2>&1; echo "ncu $v reg rc=$?"
done
timeout 120 python tools/attn_one.py > gpurun_out/attn_one.log 2>&1 && \
timeout 500 ncu --set full --clock-control none -k regex:attn_ -s 2 -c 6 -o gpurun_out/r02_attn python tools/attn_one.py > gpurun_out/attn_ncu.log 2>&1; echo "attn rc=$?"
