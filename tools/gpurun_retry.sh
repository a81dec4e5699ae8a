#!/bin/bash
# usage: tools/gpurun_retry.sh <out-file> <gpurun args...>   — retries while the pod answers "busy" (exit 3 / transient)
out=$1; shift
for i in $(seq 1 20); do
  gpurun "$@" > "$out" 2>&1
  if grep -q "status=transient\|retry in a few minutes" "$out"; then sleep 90; continue; fi
  break
done
