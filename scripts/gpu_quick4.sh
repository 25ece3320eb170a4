#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
bash scripts/gpu_stress.sh 2>&1 | cut -c1-330 | tail -6
timeout 300 python bench.py --skip-e2e --no-cpu-baseline --no-extra --steps 10 --warmup 3 2>/dev/null | cut -c1-120
