#!/bin/bash
# full GPU validation: every -m gpu test, then the default bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?" >> gpurun_out/bench2.err
tail -n 4 gpurun_out/t_all.log; tail -n 3 gpurun_out/bench2.err; head -c 600 gpurun_out/bench2.json
