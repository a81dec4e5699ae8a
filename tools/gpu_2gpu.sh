#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -x -q -m gpu > gpurun_out/t_ddp_2.log 2>&1; echo "rc=$?" >> gpurun_out/t_ddp_2.log; tail -n 3 gpurun_out/t_ddp_2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
grep '^{' gpurun_out/bench_2gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('scaling_diag'), d.get('ddp_param_drift'))"; tail -n 3 gpurun_out/bench_2gpu.err
