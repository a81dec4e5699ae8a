#!/bin/bash
# attention A/B of library variants: tests on the default build, then probe + short bench per variant
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" > gpurun_out/t_attn.log 2>&1; tail -n 2 gpurun_out/t_attn.log
for tag in base "$@" base; do
  if [ "$tag" = base ]; then unset BFK_LIB_PATH; else export BFK_LIB_PATH=$GRAFT_REPO_ROOT/bifold_b200/lib/libbifold_b200_$tag.so; fi
  echo "== $tag"
  timeout 200 python tools/attn_probe.py 2>&1 | grep fwd
  timeout 600 python bench.py --no-extra --no-cpu-baseline --no-roofline --skip-e2e --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3))"
done
