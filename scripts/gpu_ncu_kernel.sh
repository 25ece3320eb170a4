#!/bin/bash
# ncu --set full of the level-0 launch (4th build) of the kernels matching $1 (regex), $2 = how many of them run per level
K=$1; PER=${2:-1}; OUT=${3:-prof_k}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $((3*8*PER)) -c $PER -f -o gpurun_out/$OUT python bench.py --steps 1 --warmup 3 --skip-e2e --no-extra --no-cpu-baseline > gpurun_out/ncu_full_k.log 2>&1
tail -1 gpurun_out/ncu_full_k.log | cut -c1-200
