#!/bin/bash
# usage: gpu_env_ab.sh "VAR=val" "VAR2=val" ...  — each setting: short device-resident bench; first and last run = baseline
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() { env $1 timeout 600 python bench.py --no-extra --no-cpu-baseline --no-roofline --skip-e2e --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), d['launch_mode'][:24])"; }
run "BFK_NOP=1"
for s in "$@"; do run "$s"; done
run "BFK_NOP=1"
