#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=900 TAILN=8 run policy python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider
TMO=600 TAILN=3 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e"
SEC="--section SpeedOfLight --section LaunchStats --section WarpStateStats --section MemoryWorkloadAnalysis"
$B > gpurun_out/plain.log 2>&1 &&
TMO=600 TAILN=2 run ncu_gemm_fwd ncu $SEC --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -s 1 -c 4 -o gpurun_out/gemm_fwd $B
TMO=600 TAILN=2 run ncu_gemm_fus ncu $SEC --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -s 97 -c 4 -o gpurun_out/gemm_fus $B
TMO=600 TAILN=2 run ncu_gemm_bwd ncu $SEC --clock-control none --profile-from-start off -k regex:gemm_tc_kernel -s 129 -c 8 -o gpurun_out/gemm_bwd $B
TMO=600 TAILN=2 run ncu_attn ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn_fwd_tc -s 24 -c 1 -o gpurun_out/attn_fwd $B
TMO=600 TAILN=2 run ncu_attn_bwd ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn_bwd_dkv -s 0 -c 1 -o gpurun_out/attn_dkv $B
ls -la gpurun_out
