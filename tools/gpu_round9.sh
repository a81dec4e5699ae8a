#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=600 TAILN=12 run kernels python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider
TMO=900 TAILN=12 run policy python -m pytest tests/test_policy_gpu.py -m gpu -q -p no:cacheprovider
TMO=300 TAILN=2 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
BFK_GEMM_2CTA=0 TMO=300 TAILN=2 run bench_1cta python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e"
$B > gpurun_out/plain.log 2>&1 &&
TMO=900 TAILN=3 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $B
python tools/summarize_ncu.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
