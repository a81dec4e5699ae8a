#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e > gpurun_out/plain.log 2>&1 &&
timeout -k 10 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_kernel -s 1 -c 4 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e > gpurun_out/ncu_gemm.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/ncu_gemm.log; ls -la gpurun_out/
