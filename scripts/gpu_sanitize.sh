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
# round 2: heavy-parent gate split (row threshold forced down), from_mixture, overlap off
ctx3 = HemContext(0, 7)
ctx3.set_option("heavy_rows", 32); ctx3.set_option("heavy_rows_rel", 0)
ctx3.build(*g.arrays(), cluster_level=2)
ctx3.build(*g.arrays(), cluster_level=2)
m = ctx3.from_mixture(1)
print("heavy gate sizes", [ctx3.level_size(l) for l in range(3)], float(m["_rotation"].float().nan_to_num().sum()))
ctx3.set_option("overlap", 0)
ctx3.build(*c.arrays(), cluster_level=2)
print("no-overlap sizes", [ctx3.level_size(l) for l in range(3)])
# geometry-first pipeline forced on every level: pinned host inputs, bound pinned outputs (SH passes, stash, download stream)
import torch
h = synth(40000, seed=11, density=5000.0)
pin = [torch.from_numpy(x).pin_memory() for x in h.arrays()]
ctx4 = HemContext(0, 7)
ctx4.set_option("geo_first", 2)
outs = {}
for l in (1, 2, 3):
    outs[l] = [torch.empty((h.n, 3)).pin_memory(), torch.empty((h.n, 3)).pin_memory(), torch.empty(h.n).pin_memory(),
               torch.empty((h.n, 6)).pin_memory(), torch.empty((h.n, 45)).pin_memory()]
    ctx4.bind_output(l, *[b.numpy() for b in outs[l]])
ctx4.build(*[p.numpy() for p in pin], cluster_level=3)
print("geo-first sizes", [ctx4.level_size(l) for l in range(4)], float(outs[3][4][:ctx4.level_size(3)].sum()))
# ICP consumer: two small pyramids, coarse-to-fine; plus host arrays with correspondences and a fringe (no match) part
from gaussiansplattingregistration_b200.synth import rigid_copy
a = synth(20000, seed=8, density=5000.0)
b, T = rigid_copy(a, seed=9, max_rot_deg=5.0, max_trans_frac=0.02)
ca, cb = HemContext(0, 3), HemContext(0, 3)
ca.build(*a.arrays(), cluster_level=2); cb.build(*b.arrays(), cluster_level=2)
r = ca.icp_multiscale(cb, [1.0, 0.5, 0.2], [20, 10, 5])
print("icp sizes", [x.iterations for x in r], r[-1].fitness)
src = np.concatenate([a.xyz.astype(np.float64), a.xyz.astype(np.float64)[:500] + 50.0])
r2 = ca.icp(src, b.xyz.astype(np.float64), None, 0.3, max_iteration=6, with_correspondences=True)
print("icp sizes", r2.iterations, r2.n_correspondences, int((r2.correspondence < 0).sum()))
PY
for tool in memcheck racecheck initcheck; do
  echo "#### $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sizes|Invalid|hazard|Uninitialized|at .*hem" | head -20
done
