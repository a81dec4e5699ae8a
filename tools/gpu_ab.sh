#!/bin/bash
# A/B of one environment toggle on the same box: $1 = VAR, runs VAR=1, VAR=0, VAR=1, VAR=0
nproc; uptime
for v in 1 0 1 0; do
  env $1=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$1=$v', round(d['value'],1), round(d['ms_per_step'],3), 'host_enqueue', round(d['host_enqueue_ms_per_step'],2), d['clocks'])"
  uptime
done
