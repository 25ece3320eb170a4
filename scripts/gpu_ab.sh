#!/bin/bash
# A/B of an environment switch: $1 = variable name; runs the device-resident bench with it 0 and 1.
mkdir -p gpurun_out
for v in 0 1 0 1; do
  env $1=$v timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 30 > gpurun_out/ab_$v.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
print("$1=$v ms/step %.4f value %.4e"%(d['ms_per_step'], d['value']))
PY
done
