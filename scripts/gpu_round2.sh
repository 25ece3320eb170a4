#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; tail -3 gpurun_out/bench_cur.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cur.json').read().strip().splitlines()[-1])
print("value %.3e e2e %.3e ms/step %.3f e2e ms %.3f launches %d"%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches']))
print(d['kernels']); print(d['clocks']); print(d.get('cpu_baseline'))
PY
