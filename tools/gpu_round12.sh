#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=900 TAILN=6 run policy python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider
bash tools/gpu_ab.sh BFK_OVERLAP_WGRAD
