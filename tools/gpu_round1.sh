#!/bin/bash
# First GPU round: kernels one group per process (a faulting kernel must not poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
TMO=600 run k_simt python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not bf16-2 and not (bfloat16-0)" -p no:cacheprovider
TMO=300 run diag_gemm python tools/diag_gemm.py
TMO=600 run k_tc python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "bf16-2" -p no:cacheprovider
TMO=900 run policy_fp32 python -m pytest tests/test_policy_gpu.py -m gpu -q -k "fp32 or inference or generic or mask_pick" -p no:cacheprovider
TMO=900 run policy_bf16 python -m pytest tests/test_policy_gpu.py -m gpu -q -k "bf16" -p no:cacheprovider
TMO=300 run smoke python __graft_entry__.py --smoke
TMO=900 run bench python bench.py --steps 5 --warmup 3
