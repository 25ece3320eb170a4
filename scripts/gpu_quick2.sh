#!/bin/bash
# GPU tests + bench on the main library, then per-level stage timings of alternative builds given as arguments
mkdir -p gpurun_out; rm -f gpurun_out/variants.log
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --skip-e2e --no-cpu-baseline --no-extra > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-200 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
bash scripts/gpu_variants.sh "$@" 2>&1 | grep -v "^n 1000000\|^n 374" 
