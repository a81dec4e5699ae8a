#!/bin/bash
# usage: gpu_nccl_ab.sh N "ENV1=.." "ENV2=.." ...  — NCCL algorithm / channel A/B of the N-GPU training step (scaling_diag per run)
N=$1; shift
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
port=29700
for e in "BFK_NOP=1" "$@"; do
  port=$((port+7))
  env $e timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline --skip-e2e --no-roofline 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d.get('scaling_diag') or {}; print('$e', round(d['value'],1), round(d['ms_per_step'],3), 'local max', round(s.get('local_max_ms',0),3), 'comm', round(s.get('exposed_comm_share',0),4))"
done
