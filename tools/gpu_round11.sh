#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=600 TAILN=8 run kernels python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider
TMO=120 TAILN=8 run probe python tools/attn_probe.py
TMO=900 TAILN=8 run policy python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider
TMO=300 TAILN=2 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
