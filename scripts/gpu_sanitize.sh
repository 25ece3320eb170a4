#!/bin/bash
# compute-sanitizer on a small pyramid: memcheck + racecheck + initcheck
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(30000, seed=5, density=8000.0)
ctx = HemContext(0, 7)
ctx.build(*c.arrays(), cluster_level=3)
print("sizes", [ctx.level_size(l) for l in range(4)])
g = synth(3000, seed=6, density=20000.0, clip_sigma=None, scale_sigma_ln=1.0)
ctx.build(*g.arrays(), cluster_level=2)
print("sizes", [ctx.level_size(l) for l in range(3)], float(ctx.level(2)["xyz"].sum()))
# dense cloud with a fresh context: pool overflow + claim-table bail-out + redo (fast flow)
d = synth(6000, seed=7, density=200000.0)
ctx2 = HemContext(0, 7)
ctx2.build(*d.arrays(), cluster_level=2)
print("dense sizes", [ctx2.level_size(l) for l in range(3)], ctx2.step_stats(0)["n_pairs"])
PY
for tool in memcheck racecheck initcheck; do
  echo "#### $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sizes|Invalid|hazard|Uninitialized|at .*hem" | head -20
done
