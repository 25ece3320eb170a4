#!/bin/bash
# ncu launch list + full capture of the two hot kernels (level 0 of the 4th build) on the headline workload
W=${1:-3m}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 1 --warmup 3 --skip-e2e --no-extra > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
LV=8; [ "$W" = "1m" ] && LV=5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gate|k_moments" -s $((3*2*LV)) -c 2 -f -o gpurun_out/prof_hot_$W python bench.py --workload $W --steps 1 --warmup 3 --skip-e2e --no-extra > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
