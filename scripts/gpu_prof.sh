#!/bin/bash
# bench + ncu launch list + full capture of the two hot kernels (level 0 of the 3rd build)
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; tail -c 600 gpurun_out/bench_cur.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 260 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 2 --warmup 3 --skip-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gate|k_moments" -s 30 -c 2 -f -o gpurun_out/prof_cur python bench.py --steps 1 --warmup 3 --skip-e2e > gpurun_out/ncu_full_cur.log 2>&1
tail -2 gpurun_out/ncu_full_cur.log
