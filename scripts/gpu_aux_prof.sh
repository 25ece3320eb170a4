#!/bin/bash
# ncu --set full of the level-0 launches (4th build) of the kernels AROUND the two hot ones
mkdir -p gpurun_out
K="k_chunk_copy|k_sumlw_claims|k_sorted_gather|k_finish|k_orphans|k_cellkey|k_boxhist|k_import|k_prep"
# per build: level 0 runs all 9, levels 1..7 run 6 each (no import / prep / boxhist): 51 launches; skip 3 warm-up builds
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $((3*51)) -c 9 -f -o gpurun_out/prof_aux python bench.py --steps 1 --warmup 3 --skip-e2e --no-extra --no-cpu-baseline > gpurun_out/ncu_full_aux.log 2>&1
tail -1 gpurun_out/ncu_full_aux.log | cut -c1-200
