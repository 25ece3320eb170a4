#!/bin/bash
# GPU tests, a short bench (device-resident value only) and the ncu launch list of the same command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --skip-e2e --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-300 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_quick.csv python bench.py --steps 1 --warmup 3 --skip-e2e --no-extra --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
