#!/bin/bash
# does host CPU contention slow the step?  run the bench alone, then with 32 busy-loop processes competing for the 16 cores
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline --skip-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$1', round(d['value'],1), round(d['ms_per_step'],3), 'host_enqueue', round(d['host_enqueue_ms_per_step'],2))"; }
run alone
pids=""
for i in $(seq 32); do python -c "
import time
t=time.time()
while time.time()-t<40: pass" & pids="$pids $!"; done
sleep 1
run loaded
for p in $pids; do kill $p 2>/dev/null; done
wait 2>/dev/null
run alone_again
