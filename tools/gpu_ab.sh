#!/bin/bash
# A/B of library variants built with BFK_BUILD_TAG (tools: bifold_b200/build.py): usage gpu_ab.sh tag1 tag2 ...
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for tag in base "$@"; do
  if [ "$tag" = base ]; then unset BFK_LIB_PATH; else export BFK_LIB_PATH=$GRAFT_REPO_ROOT/bifold_b200/lib/libbifold_b200_$tag.so; fi
  echo "== $tag"
  timeout 300 python tools/bench_epilogue.py 2>&1 | tail -n 3
  timeout 600 python bench.py --no-extra --no-cpu-baseline --no-roofline --skip-e2e --steps 20 2> gpurun_out/ab_$tag.err | head -c 200; echo
done > gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
