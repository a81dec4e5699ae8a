#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python tools/bench_epilogue.py > gpurun_out/epilogue2.log 2>&1; cat gpurun_out/epilogue2.log
for v in gg plain; do
  timeout 120 python tools/gemm_one.py $v > /dev/null 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/gemm_$v python tools/gemm_one.py $v > gpurun_out/ncu_gemm_$v.log 2>&1
  echo "ncu $v rc=$?"
done
