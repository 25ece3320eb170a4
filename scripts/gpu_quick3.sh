#!/bin/bash
# GPU tests + a short device-resident bench (no profiler passes)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --skip-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-200 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
python scripts/gpu_base3m.py 2>&1 | grep -E "step [0-7]:|pyramid" | sed -E 's/.*n_in ([0-9]+).*ms tot ([0-9.]+) prep ([0-9.]+) grid ([0-9.]+) gate ([0-9.]+) \(k ([0-9.]+)\).*mom ([0-9.]+) \(k ([0-9.]+)\).*/n \1 tot \2 prep \3 grid \4 gate \5 k_gate \6 mom \7 k_mom \8/' | head -9
