#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=300 TAILN=60 run attn_tc python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention and bfloat16-0" -p no:cacheprovider -x
TMO=900 run policy_bf16 python -m pytest tests/test_policy_gpu.py -m gpu -q -k "bf16" -p no:cacheprovider
TMO=300 run smoke python __graft_entry__.py --smoke
TMO=900 run bench python bench.py --steps 5 --warmup 3
