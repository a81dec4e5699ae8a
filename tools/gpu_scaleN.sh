#!/bin/bash
# usage: gpu_scaleN.sh N   — config-2 bench line at N GPUs (with scaling_diag)
N=${1:-2}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-roofline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
grep '^{' gpurun_out/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d.get('scaling_diag') or {}; print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), [round(x,2) for x in s.get('local_ms_per_rank',[])], 'straggler', round(s.get('straggler_share',0),4), 'comm', round(s.get('exposed_comm_share',0),4))"
