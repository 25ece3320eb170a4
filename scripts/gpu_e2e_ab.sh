#!/bin/bash
# geometry-first test + the e2e leg of bench.py with the pipeline on and off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "geometry_first or bound_outputs" 2>&1 | tail -5
for G in 1 0; do
  HEM_GEO_FIRST=$G timeout 600 python bench.py --no-cpu-baseline --no-extra --steps 5 --warmup 3 > gpurun_out/bench_e2e_$G.json 2> gpurun_out/bench_e2e_$G.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_e2e_$G.json').read().strip().splitlines()[-1])
print('geo_first=$G value %.3e (%.3f ms)  e2e %.3e (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
PY
  tail -2 gpurun_out/bench_e2e_$G.err
done
