#!/bin/bash
# round-2 state: every -m gpu test, the default bench line, then the ncu launch list of a short bench run
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?" >> gpurun_out/bench_r02_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02_final.csv \
  python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline --skip-e2e --no-roofline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?" >> gpurun_out/ncu_bench.log
tail -n 4 gpurun_out/t_all.log; tail -n 3 gpurun_out/bench_r02_final.err; head -c 3000 gpurun_out/bench_r02_final.json; tail -n 2 gpurun_out/ncu_bench.log
