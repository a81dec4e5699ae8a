#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=300 TAILN=40 run attn_tc python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e > gpurun_out/plain.log 2>&1 &&
TMO=900 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e
python tools/summarize_ncu.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -60 gpurun_out/launch_summary.txt
