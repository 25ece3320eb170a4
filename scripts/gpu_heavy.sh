#!/bin/bash
# heavy-parent gate split: parity tests, then timings (ordinary pyramids + the unclipped heavy-tail stress cloud)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "heavy or hub or golden or free_running or properties" 2>&1 | tail -5 | tee gpurun_out/heavy_tests.log
timeout 300 python scripts/gpu_base3m.py 2>&1 | grep -E "step|pyramid" | cut -c1-260 | tee gpurun_out/heavy_time.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/heavy_stress.log
import sys, time; sys.path.insert(0, '.')
import torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(1_000_000, seed=1, density=5000, scale_sigma_ln=1.0, clip_sigma=None)
dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
for hr in (0, 128):
    ctx = HemContext(0, 7); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_option("timing", 1); ctx.set_option("heavy_rows", hr)
    for r in range(3):
        torch.cuda.synchronize(); t = time.time(); ctx.build(*dev, cluster_level=3); torch.cuda.synchronize(); dt = time.time() - t
    print(f"stress sigma_ln=1.0 unclipped 1M x 3 levels, heavy_rows={hr}: wall {dt*1e3:.2f} ms", [(ctx.step_stats(l)['n_in'], round(ctx.step_stats(l)['ms_k_gate'], 3), round(ctx.step_stats(l)['ms_k_moments'], 3), round(ctx.step_stats(l)['ms_total'], 3)) for l in range(3)])
    ctx.close()
PY
