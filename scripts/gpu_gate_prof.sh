#!/bin/bash
# ncu --set full of the gate kernels on level 0 of the 1M cloud
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(1_000_000, seed=1, density=5000)
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
ctx = HemContext(0, 7); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_option("timing", 1)
for r in range(3):
    ctx.build(*dev, cluster_level=1)
s = ctx.step_stats(0)
print({k: s[k] for k in ("n_parents", "n_pairs", "n_wide_parents", "n_tile_fallbacks", "tile_halo", "ms_k_gate", "ms_gate", "ms_total")})
PY
timeout 120 python /tmp/one.py 2>&1 | tail -2 | tee gpurun_out/gate_one.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gate" -s 4 -c 2 -f -o gpurun_out/prof_gate_r2 python /tmp/one.py > gpurun_out/ncu_gate_r2.log 2>&1
tail -3 gpurun_out/ncu_gate_r2.log
