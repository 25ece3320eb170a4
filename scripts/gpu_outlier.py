import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
for far in (0.0, 1e2, 1e3, 1e4):
    c = synth(1_000_000, seed=1, density=5000.0)
    if far > 0:
        rng = np.random.default_rng(3)
        idx = rng.choice(c.n, 20, replace=False)
        c.xyz[idx] += (rng.standard_normal((20, 3)) * far).astype(np.float32)
    dev = [torch.from_numpy(a).cuda() for a in c.arrays()]
    ctx = HemContext(0, 7); ctx.set_option("timing", 1)
    for r in range(2):
        torch.cuda.synchronize(); t=time.perf_counter(); ctx.build(*dev, cluster_level=3); dt=time.perf_counter()-t
    s=ctx.step_stats(0)
    print(f"outliers at ~{far:g}: build {dt*1e3:.2f} ms; level 0: cells {s['grid_cells']} h {s['cell_size']:.3f} gate {s['ms_k_gate']:.3f} ms; sizes {[ctx.level_size(l) for l in range(4)]}")
