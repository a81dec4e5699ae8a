#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 python tools/attn_trace_bwd.py > gpurun_out/attn_trace_bwd.log 2>&1

