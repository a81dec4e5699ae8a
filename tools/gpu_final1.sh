#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-25} gpurun_out/$name.log; }
TMO=300 TAILN=4 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=600 TAILN=1 run bench_full python bench.py --steps 20 --warmup 5
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --skip-e2e"
TMO=900 TAILN=2 run ncu_launches ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv $B
python tools/summarize_ncu.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
timeout 60 python tools/bench_gemm_parts.py > gpurun_out/parts_plain.log 2>&1 &&
TMO=400 TAILN=2 run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 2 -c 3 -o gpurun_out/gemm_fc1 python tools/bench_gemm_parts.py
