#!/bin/bash
# round-2 GPU run 1: new tests first (fast feedback), then the whole GPU suite, then the default bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_input_stage.py tests/test_checkpoint_gpu.py tests/test_ops_cpu.py -x -q -m gpu > gpurun_out/t_new.log 2>&1; echo "new rc=$?" >> gpurun_out/t_new.log
timeout 1200 python -m pytest tests/test_policy_gpu.py -x -q -m gpu -s > gpurun_out/t_policy.log 2>&1; echo "policy rc=$?" >> gpurun_out/t_policy.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/t_kernels.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/bench1.err
for f in gpurun_out/t_new.log gpurun_out/t_policy.log gpurun_out/t_kernels.log; do tail -n 3 $f; done; tail -n 5 gpurun_out/bench1.err; head -c 3000 gpurun_out/bench1.json
