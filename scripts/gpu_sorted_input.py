"""Experiment: does the moment kernel's gather suffer from the (random) canonical row order?  Same 3M cloud, input rows
in generator order vs pre-sorted by a coarse spatial key."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaussiansplattingregistration_b200.synth import synth
from gaussiansplattingregistration_b200.hem import HemContext
c = synth(3_000_000, seed=1, density=5000, aspect=(2, 1, 0.5))
for mode in ("generator order", "sorted by 0.3-unit cells (z,y,x)"):
    arrs = list(c.arrays())
    if mode != "generator order":
        k = np.floor(c.xyz / 0.3).astype(np.int64)
        order = np.lexsort((k[:, 0], k[:, 1], k[:, 2]))
        arrs = [np.ascontiguousarray(a[order]) for a in arrs]
    dev = [torch.from_numpy(a).cuda() for a in arrs]
    ctx = HemContext(0, 7); ctx.set_stream(torch.cuda.current_stream().cuda_stream); ctx.set_option("timing", 1)
    for r in range(3):
        ctx.build(*dev, cluster_level=3)
    for l in range(3):
        s = ctx.step_stats(l)
        print(f"{mode}: level {l} n {s['n_in']} tot {s['ms_total']:.3f} grid {s['ms_grid']:.3f} k_gate {s['ms_k_gate']:.3f} sumlw {s['ms_sumlw']:.3f} k_mom {s['ms_k_moments']:.3f}")
    ctx.close(); del dev
