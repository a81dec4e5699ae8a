#!/bin/bash
# usage: gpu_bucket_ab.sh N mb1 mb2 ...   — DDP bucket-size A/B at N GPUs (BFK_DDP_BUCKET_MB)
N=$1; shift
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
port=29600
for mb in "$@"; do
  port=$((port+7))
  BFK_DDP_BUCKET_MB=$mb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline --skip-e2e --no-roofline > gpurun_out/bucket_${N}_$mb.json 2> gpurun_out/bucket_${N}_$mb.err
  echo "== bucket $mb MB rc=$?"
  grep '^{' gpurun_out/bucket_${N}_$mb.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d.get('scaling_diag') or {}; print(round(d['value'],1), round(d['ms_per_step'],3), [round(x,2) for x in s.get('local_ms_per_rank',[])], 'straggler', round(s.get('straggler_share',0),4), 'comm', round(s.get('exposed_comm_share',0),4))"
done
