#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_2gpu.log | tail -3
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-batch 2 > gpurun_out/bench_ref_2gpu.log 2>&1; echo "exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_ref_2gpu.log | tail -2
