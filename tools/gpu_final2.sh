#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>/dev/null; echo "exit $?"; tail -n 1 gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TMO=600 run kernels python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider
TMO=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=600 CUT=3000 run bench_full python bench.py --steps 20 --warmup 5
TMO=300 run bench_infer python bench.py --mode infer --steps 20 --warmup 5 --no-cpu-baseline --no-roofline
TMO=300 run bench_ctx python bench.py --context 3 --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
TMO=300 run bench_384 python bench.py --image-size 384 --batch 32 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
TMO=300 run bench_ref python bench.py --impl reference --steps 1 --warmup 0
