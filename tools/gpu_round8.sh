#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=900 TAILN=12 run policy python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider
TMO=300 TAILN=2 run bench_overlap python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
BFK_OVERLAP_TOWERS=0 TMO=300 TAILN=2 run bench_serial python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
