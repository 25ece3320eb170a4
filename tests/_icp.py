"""TEST INFRASTRUCTURE — point-to-point ICP stand-in with Open3D's definitions.

Open3D (pinned 0.16.0 by the reference's requirements.txt) is not in this image, so the
downstream check of SURVEY.md §8c/§8d ("ICP fitness/RMSE on the produced pyramids within 1e-3")
runs the SAME stand-in on the oracle pyramid and on the GPU pyramid:
  correspondences = nearest target point of each transformed source point within max_correspondence
  fitness = #correspondences / #source points, inlier_rmse = sqrt(sum d^2 / #correspondences)
  update = rigid Kabsch/SVD (no scale); stop when |d fitness| < rel_fitness and |d rmse| < rel_rmse
  or at max_iteration (reference src/utils/local_registration_util.py:54-55,86-98).
Coarse-to-fine driver as the reference intends it (qt_multiscale_registrator.py:203-230):
coarsest level first, each level initialised with the previous transformation.
"""
import numpy as np
from scipy.spatial import cKDTree


def _evaluate(src, tree, T, max_corr):
    p = src @ T[:3, :3].T + T[:3, 3]
    d, idx = tree.query(p, k=1, distance_upper_bound=max_corr)
    ok = np.isfinite(d)
    n = int(ok.sum())
    fitness = n / max(len(src), 1)
    rmse = float(np.sqrt((d[ok] ** 2).sum() / n)) if n else 0.0
    return p, idx, ok, fitness, rmse


def icp_point_to_point(src, tgt, init, max_corr, max_iter, rel_fitness=1e-6, rel_rmse=1e-6):
    src = np.asarray(src, np.float64)
    tgt = np.asarray(tgt, np.float64)
    tree = cKDTree(tgt)
    T = np.array(init, np.float64)
    p, idx, ok, fitness, rmse = _evaluate(src, tree, T, max_corr)
    for _ in range(max_iter):
        if ok.sum() < 3:
            break
        a, b = p[ok], tgt[idx[ok]]
        ca, cb = a.mean(0), b.mean(0)
        H = (a - ca).T @ (b - cb)
        U, _, Vt = np.linalg.svd(H)
        D = np.diag([1.0, 1.0, np.sign(np.linalg.det(Vt.T @ U.T))])
        R = Vt.T @ D @ U.T
        t = cb - R @ ca
        U4 = np.eye(4)
        U4[:3, :3], U4[:3, 3] = R, t
        T = U4 @ T
        p, idx, ok, f2, r2 = _evaluate(src, tree, T, max_corr)
        done = abs(f2 - fitness) < rel_fitness and abs(r2 - rmse) < rel_rmse
        fitness, rmse = f2, r2
        if done:
            break
    return T, fitness, rmse


def multiscale_icp(pyr_src, pyr_tgt, max_corrs, max_iters):
    """pyr_* = [level 0 (full), level 1, ...]; coarsest first, chained transformations."""
    assert len(pyr_src) == len(pyr_tgt) == len(max_corrs) == len(max_iters)
    T = np.eye(4)
    out = []
    for i in range(len(pyr_src)):
        s, t = pyr_src[-(i + 1)], pyr_tgt[-(i + 1)]
        T, fit, rmse = icp_point_to_point(s, t, T, max_corrs[i], max_iters[i])
        out.append((fit, rmse))
    return T, out
