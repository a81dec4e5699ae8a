#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1; echo "exit $?"; tail -5 gpurun_out/bench_2gpu.log
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "exit $?"; tail -8 gpurun_out/ddp_check.log
