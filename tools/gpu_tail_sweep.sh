#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
: > gpurun_out/tail_sweep.log
for hm in 2 4 8; do for minb in 3 4; do
  echo "== BFK_HM_BANDS=$hm MINB=$minb" >> gpurun_out/tail_sweep.log
  BFK_HM_BANDS=$hm BFK_HM_MINB=$minb timeout 200 python tools/bench_tail.py 2>&1 | grep -E "^heatmap_fwd" >> gpurun_out/tail_sweep.log
done; done
for hb in 1 2 4; do
  echo "== BFK_HB_BANDS=$hb" >> gpurun_out/tail_sweep.log
  BFK_HB_BANDS=$hb timeout 200 python tools/bench_tail.py 2>&1 | grep -E "^heatmap_bwd" >> gpurun_out/tail_sweep.log
done
cat gpurun_out/tail_sweep.log
