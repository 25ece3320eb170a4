#!/bin/bash
# the unclipped heavy-tail stress cloud (sigma_ln = 1.0, 1 M splats, 3 levels) with the default settings
mkdir -p gpurun_out
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/heavy_stress.log
import sys, time; sys.path.insert(0, '.')
import torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(1_000_000, seed=1, density=5000, scale_sigma_ln=1.0, clip_sigma=None)
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
ctx = HemContext(0, 7); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_option("timing", 1)
for r in range(4):
    torch.cuda.synchronize(); t = time.time(); ctx.build(*dev, cluster_level=3); torch.cuda.synchronize(); dt = time.time() - t
    print(f"stress sigma_ln=1.0 unclipped 1M x 3 levels, defaults, build {r}: wall {dt*1e3:.2f} ms", [(ctx.step_stats(l)['n_in'], round(ctx.step_stats(l)['ms_k_gate'], 3), round(ctx.step_stats(l)['ms_k_moments'], 3), round(ctx.step_stats(l)['ms_total'], 3)) for l in range(3)])
    if r == 3:
        st = ctx.step_stats(0)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
        print("launches", ctx.info("launches"))
ctx.close()
PY
