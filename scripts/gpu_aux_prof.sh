#!/bin/bash
# tests + bench, then ncu: launch list and a full capture of the level-0 launches of the kernels AROUND the two hot ones
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --skip-e2e --no-cpu-baseline > gpurun_out/bench_aux.json 2> gpurun_out/bench_aux.err; cut -c1-400 gpurun_out/bench_aux.json; tail -3 gpurun_out/bench_aux.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_aux.json').read().strip().splitlines()[-1])
print([round(l['ms'],3) for l in d['levels']], d['workloads']['1m']['level_ms'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_aux.csv python bench.py --steps 1 --warmup 3 --skip-e2e --no-extra --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
K="k_chunk_copy|k_sumlw_claims|k_sorted_gather|k_ptab|k_finish|k_orphans|k_cellkey|k_boxhist"
# level 0 of the 4th build: 8 kernels x 8 levels x 3 warm-up builds skipped, the next 8 launches are level 0's
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $((3*8*8 - 3*7)) -c 8 -f -o gpurun_out/prof_aux python bench.py --steps 1 --warmup 3 --skip-e2e --no-extra --no-cpu-baseline > gpurun_out/ncu_full_aux.log 2>&1
tail -1 gpurun_out/ncu_full_aux.log | cut -c1-200
