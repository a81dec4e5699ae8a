#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python tools/attn_probe.py > gpurun_out/attn_probe_default.log 2>&1
BFK_ATTN_V2_FWD=1 timeout 200 python tools/attn_probe.py > gpurun_out/attn_probe_v2fwd.log 2>&1
BFK_ATTN_V2=0 timeout 200 python tools/attn_probe.py > gpurun_out/attn_probe_v1.log 2>&1
timeout 300 python tools/bench_tail.py > gpurun_out/tail_bench.log 2>&1
for f in attn_probe_default attn_probe_v2fwd attn_probe_v1 tail_bench; do echo "== $f"; head -n 14 gpurun_out/$f.log; done
