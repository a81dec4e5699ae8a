#!/bin/bash
# usage: gpu_ddp_n.sh N
mkdir -p gpurun_out
N=$1
date +%s > /tmp/t0
timeout -k 5 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_${N}gpu.log 2>&1; echo "exit $? after $(( $(date +%s) - $(cat /tmp/t0) )) s"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_${N}gpu.log | grep "^{" | tail -1 | cut -c1-900; nproc
