#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python tools/bench_epilogue.py > gpurun_out/epi.log 2>&1; cat gpurun_out/epi.log
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e"
$B > gpurun_out/plain.log 2>&1 &&
timeout -k 10 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -s 157 -c 1 -o gpurun_out/gemm_dgelu $B > gpurun_out/ncu_dgelu.log 2>&1; tail -2 gpurun_out/ncu_dgelu.log
