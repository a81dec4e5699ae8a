#!/bin/bash
mkdir -p gpurun_out
date +%s > /tmp/t0
timeout -k 5 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "exit $? after $(( $(date +%s) - $(cat /tmp/t0) )) s"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_2gpu.log | grep "^{" | tail -1 | cut -c1-1500
