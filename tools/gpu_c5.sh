#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --tower so400m --image-size 384 --batch 16 --steps 5 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "rc=$?"
timeout 600 python bench.py --context 3 --batch 16 --steps 5 --no-extra --no-cpu-baseline --skip-e2e > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"
head -c 300 gpurun_out/bench_c5.json; echo; head -c 300 gpurun_out/bench_c3.json
