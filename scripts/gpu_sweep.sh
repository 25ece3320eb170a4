#!/bin/bash
# sweep of one environment variable over the device-resident bench: $1 = name, rest = values
mkdir -p gpurun_out
name=$1; shift
for v in "$@" "$1"; do
  env $name=$v timeout 300 python bench.py --no-cpu-baseline --skip-e2e --steps 30 > gpurun_out/sw.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/sw.json').read().strip().splitlines()[-1])
print("$name=$v ms/step %.4f value %.4e"%(d['ms_per_step'], d['value']), d.get('kernels'))
PY
done
