#!/bin/bash
# one iteration: kernel + policy tests, epilogue / LN probes, short bench
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_policy_gpu.py -x -q -m gpu > gpurun_out/t_iter.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_iter.log
timeout 200 python tools/ln_probe.py > gpurun_out/ln_probe.log 2>&1
BFK_LORA_DOWN_FUSED=1 timeout 200 python tools/ln_probe.py > gpurun_out/ln_probe_fused.log 2>&1
timeout 300 python tools/bench_epilogue.py > gpurun_out/epilogue.log 2>&1
timeout 600 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?" >> gpurun_out/bench_iter.err
tail -n 4 gpurun_out/t_iter.log; cat gpurun_out/ln_probe.log gpurun_out/ln_probe_fused.log gpurun_out/epilogue.log; tail -n 2 gpurun_out/bench_iter.err; head -c 400 gpurun_out/bench_iter.json
