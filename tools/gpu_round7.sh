#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=900 TAILN=12 run policy python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider
TMO=600 TAILN=3 run bench_train python bench.py --steps 10 --warmup 3
TMO=600 TAILN=3 run bench_infer python bench.py --mode infer --steps 10 --warmup 3
TMO=600 TAILN=3 run bench_ctx python bench.py --context 3 --batch 16 --steps 5 --warmup 3 --no-cpu-baseline
TMO=600 TAILN=3 run bench_384 python bench.py --image-size 384 --batch 32 --steps 5 --warmup 3 --no-cpu-baseline
