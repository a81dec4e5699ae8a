#!/bin/bash
# N-GPU runs: DDP tests, default line (config 2) and config 3 (context T=3), both with scaling_diag.  usage: gpu_scale8.sh [N]
N=${1:-8}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_ddp_gpu.py -x -q -m gpu > gpurun_out/t_ddp_$N.log 2>&1; echo "ddp tests rc=$?"; tail -n 2 gpurun_out/t_ddp_$N.log
timeout 600 $TR --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
timeout 600 $TR --master-port 29547 bench.py --gpus $N --steps 20 --warmup 5 --context 3 --batch 16 --no-extra --no-cpu-baseline --no-roofline > gpurun_out/bench_${N}gpu_c3.json 2> gpurun_out/bench_${N}gpu_c3.err; echo "rc=$?"
timeout 600 $TR --master-port 29557 bench.py --gpus $N --steps 20 --warmup 5 --mode infer --no-extra --no-cpu-baseline --no-roofline > gpurun_out/bench_${N}gpu_infer.json 2> gpurun_out/bench_${N}gpu_infer.err; echo "rc=$?"
for f in bench_${N}gpu bench_${N}gpu_c3 bench_${N}gpu_infer; do echo "== $f"; grep '^{' gpurun_out/$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('scaling_diag'), d.get('ddp_param_drift'))"; done
