"""Config 5 of BASELINE.json in full: two 1 M-splat clouds (one a rigid copy of the other), 3-level pyramids built on the
GPU, coarse-to-fine point-to-point ICP over them — on the GPU (hem_icp_multiscale) and with the CPU oracle (scipy cKDTree,
all host threads) on the same levels.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
from gaussiansplattingregistration_b200.hem import HemContext  # noqa: E402
from gaussiansplattingregistration_b200.synth import rigid_copy, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cpu = "--no-cpu" not in sys.argv
L = 3
a = synth(n, seed=501, density=5000.0)
b, T_true = rigid_copy(a, seed=502, max_rot_deg=6.0, max_trans_frac=0.02)
spacing = (1.0 / 5000.0) ** (1.0 / 3.0)
corr = [15 * spacing, 7.5 * spacing, 6 * spacing, 3 * spacing]
iters = [50, 30, 20, 10]
out = dict(n=n, levels=L, max_correspondence=corr, max_iteration=iters)
with HemContext(0, seed=13) as ca, HemContext(0, seed=13) as cb:
    ca.build(*a.arrays(), cluster_level=L)
    cb.build(*b.arrays(), cluster_level=L)
    sizes = [ca.level_size(l) for l in range(L + 1)]
    # BASELINE configs[4] end to end with inputs resident in HBM: two pyramids + the coarse-to-fine registration
    import torch
    da = [torch.from_numpy(x).cuda() for x in a.arrays()]
    db = [torch.from_numpy(x).cuda() for x in b.arrays()]
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ca.build(*da, cluster_level=L)
        cb.build(*db, cluster_level=L)
        ca.icp_multiscale(cb, corr, iters)
        pipe = time.perf_counter() - t0
    out["pipeline_ms_wall"] = pipe * 1e3
    # the two pyramids are independent: built from two host threads (ctypes drops the GIL) on the two contexts'
    # streams, their launch-bound small levels overlap on the one GPU
    import threading
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=c.build, args=tuple(d), kwargs=dict(cluster_level=L)) for c, d in ((ca, da), (cb, db))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        both = time.perf_counter() - t0
        ca.icp_multiscale(cb, corr, iters)
        pipe2 = time.perf_counter() - t0
    out["two_builds_concurrent_ms_wall"] = both * 1e3
    out["pipeline_concurrent_ms_wall"] = pipe2 * 1e3
    for rep in range(3):
        t0 = time.perf_counter()
        res = ca.icp_multiscale(cb, corr, iters)
        wall = time.perf_counter() - t0
    out["sizes"] = sizes
    out["gpu_ms_wall"] = wall * 1e3
    out["gpu_scales"] = [dict(level=L - i, fitness=r.fitness, rmse=r.inlier_rmse, iterations=r.iterations, ms=r.ms_device,
                              cells=r.grid_cells) for i, r in enumerate(res)]
    T = res[-1].transformation
    out["err_R"] = float(np.abs(T[:3, :3] - T_true[:3, :3]).max())
    out["err_t"] = float(np.abs(T[:3, 3] - T_true[:3, 3]).max())
    if cpu:
        import icp_oracle
        pa = [ca.level(l)["xyz"].astype(np.float64) for l in range(L + 1)]
        pb = [cb.level(l)["xyz"].astype(np.float64) for l in range(L + 1)]
        t0 = time.perf_counter()
        ref = icp_oracle.multiscale(pa, pb, corr, iters)
        out["cpu_ms_wall"] = (time.perf_counter() - t0) * 1e3
        out["cpu_scales"] = [dict(fitness=r["fitness"], rmse=r["inlier_rmse"], iterations=r["iterations"]) for r in ref]
        out["max_dT"] = float(max(np.abs(g.transformation - r["transformation"]).max() for g, r in zip(res, ref)))
print(json.dumps(out))
