#!/bin/bash
# One GPU-box visit: tests, smoke, bench, ncu launch list + full captures of the two hot kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gate|k_moments" -s 20 -c 2 -f -o gpurun_out/prof_hot python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
