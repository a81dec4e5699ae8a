#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 tools/ubench/pipes > gpurun_out/ubench_pipes.log 2>&1; cat gpurun_out/ubench_pipes.log
