#!/bin/bash
# 8-GPU runs: default line, NCCL CTA-limit A/B, config 3 (context T=3).  usage: gpu_scale8.sh [N]
N=${1:-8}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
NCCL_MAX_CTAS=4 timeout 600 $TR --master-port 29527 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline --skip-e2e --no-roofline > gpurun_out/bench_${N}gpu_ctas4.json 2> gpurun_out/bench_${N}gpu_ctas4.err; echo "rc=$?"
timeout 600 $TR --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 --context 3 --batch 16 --no-extra --no-cpu-baseline > gpurun_out/bench_${N}gpu_c3.json 2> gpurun_out/bench_${N}gpu_c3.err; echo "rc=$?"
for f in bench_${N}gpu bench_${N}gpu_ctas4 bench_${N}gpu_c3; do echo "== $f"; grep '^{' gpurun_out/$f.json | head -c 330; echo; tail -n 2 gpurun_out/$f.err; done
