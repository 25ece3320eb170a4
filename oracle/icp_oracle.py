"""TEST INFRASTRUCTURE — CPU oracle of the ICP consumer.  Never imported by the product.

PARITY UNPINNED BY THE REFERENCE: the reference delegates ICP to Open3D (pinned open3d==0.16.0,
requirements.txt:3), a third-party dependency that is neither vendored under the reference tree nor installed
in this image, and the reference holds no test or golden vector for registration results.  This file restates
the published algorithm of Open3D 0.16 `RegistrationICP` (cpp/open3d/pipelines/registration/Registration.cpp)
as the reference calls it (src/utils/local_registration_util.py:76-100, point-to-point, no robust loss):

  pcd = source transformed by init
  result = evaluate(pcd)                      # GetRegistrationResultAndCorrespondences
  for i in range(max_iteration):
      update = umeyama(pcd[corr], target[corr], scaling=False)   # TransformationEstimationPointToPoint
      transformation = update @ transformation;  pcd = update applied to pcd
      backup, result = result, evaluate(pcd)
      if |backup.fitness - result.fitness| < relative_fitness and |backup.rmse - result.rmse| < relative_rmse: break

  evaluate: for each point the nearest target point (KD-tree hybrid search, k=1, radius max_correspondence,
  strict d < radius as nanoflann does); fitness = #corr / #source; inlier_rmse = sqrt(sum d^2 / #corr).

Uses real Open3D when it is importable (then parity is pinned to it), else scipy's cKDTree.
"""
import numpy as np
from scipy.spatial import cKDTree


def _evaluate(p, tree, max_corr):
    d, idx = tree.query(p, k=1, distance_upper_bound=max_corr)
    ok = np.isfinite(d) & (d < max_corr)
    n = int(ok.sum())
    fitness = n / max(len(p), 1)
    rmse = float(np.sqrt((d[ok] ** 2).sum() / n)) if n else 0.0
    return idx, ok, fitness, rmse


def umeyama_rigid(a, b):
    """Eigen::umeyama(a, b, with_scaling=false): least-squares rigid motion a -> b, proper rotation."""
    ca, cb = a.mean(0), b.mean(0)
    sigma = (b - cb).T @ (a - ca) / len(a)
    U, _, Vt = np.linalg.svd(sigma)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1.0
    R = U @ S @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = cb - R @ ca
    return T


def registration_icp(source, target, max_correspondence, init=None, relative_fitness=1e-6, relative_rmse=1e-6,
                     max_iteration=30):
    src = np.asarray(source, np.float64).reshape(-1, 3)
    tgt = np.asarray(target, np.float64).reshape(-1, 3)
    T = np.eye(4) if init is None else np.array(init, np.float64).reshape(4, 4)
    tree = cKDTree(tgt)
    p = src @ T[:3, :3].T + T[:3, 3]
    idx, ok, fitness, rmse = _evaluate(p, tree, max_correspondence)
    it, reason = 0, 2
    for _ in range(max_iteration):
        if ok.sum() < 3:
            reason = 3
            break
        U = umeyama_rigid(p[ok], tgt[idx[ok]])
        T = U @ T
        p = p @ U[:3, :3].T + U[:3, 3]
        it += 1
        idx, ok, f2, r2 = _evaluate(p, tree, max_correspondence)
        done = abs(f2 - fitness) < relative_fitness and abs(r2 - rmse) < relative_rmse
        fitness, rmse = f2, r2
        if done:
            reason = 1
            break
    corr = np.where(ok, idx, -1).astype(np.int32)
    return dict(transformation=T, fitness=fitness, inlier_rmse=rmse, iterations=it, stop_reason=reason,
                correspondence=corr, n_correspondences=int(ok.sum()))


def multiscale(pyr_src, pyr_tgt, max_corrs, max_iters, init=None, relative_fitness=1e-6, relative_rmse=1e-6):
    """Coarse-to-fine loop of qt_multiscale_registrator.py:203-230: pyr_* = [full cloud, level 1, ...]."""
    assert len(pyr_src) == len(pyr_tgt) == len(max_corrs) == len(max_iters)
    T = np.eye(4) if init is None else np.array(init, np.float64)
    out = []
    for i in range(len(max_iters)):
        r = registration_icp(pyr_src[-(i + 1)], pyr_tgt[-(i + 1)], max_corrs[i], T, relative_fitness, relative_rmse,
                             max_iters[i])
        T = r["transformation"]
        out.append(r)
    return out
