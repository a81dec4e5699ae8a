#!/bin/bash
# usage: gpu_multi.sh N   — DDP tests (N >= 2) and the bench at N GPUs
N=${1:-2}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddp_gpu.py tests/test_policy_gpu.py -x -q -m gpu -k "ddp or two_rank or lora_dropout_masks" > gpurun_out/t_ddp_$N.log 2>&1; echo "rc=$?" >> gpurun_out/t_ddp_$N.log; tail -n 3 gpurun_out/t_ddp_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/bench_${N}gpu.err
tail -n 3 gpurun_out/bench_${N}gpu.err; grep '^{' gpurun_out/bench_${N}gpu.json | head -c 400
