#!/bin/bash
export BFK_GEMM_2CTA=1
mkdir -p gpurun_out
timeout -k 10 100 python tools/bench_epilogue.py 2>&1 | head -3
timeout -k 10 100 python tools/bench_epilogue.py > gpurun_out/plain.log 2>&1 &&
timeout -k 10 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/gemm2cta python tools/bench_epilogue.py > gpurun_out/ncu2cta.log 2>&1; tail -2 gpurun_out/ncu2cta.log
