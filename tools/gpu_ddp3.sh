#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/ddp_check.log | tail -4
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_2gpu.log | tail -2 | cut -c1-1800
