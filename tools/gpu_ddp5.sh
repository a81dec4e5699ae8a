#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/ddp_graph_check.py > gpurun_out/ddp_graph_check.log 2>&1; echo "exit $?"; grep -v "^\*\|OMP_NUM" gpurun_out/ddp_graph_check.log | tail -4 | cut -c1-300
bash tools/gpu_ddp4.sh | cut -c1-420
