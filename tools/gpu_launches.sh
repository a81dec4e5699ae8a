#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e"
$B > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
python tools/summarize_ncu.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -48 gpurun_out/launch_summary.txt
