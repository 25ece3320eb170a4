"""Second caller of the HEM boundary: plane-inlier merging.

Mirror of the array logic of the reference's ``PlaneInlierMergingWorker``
(src/gui/workers/downsampling/qt_plane_merging.py:86-177) without the Qt shell: the inliers of every
fitted plane are reduced by their own HEM pyramid (``process_all_planes`` :132-152), the splats that
belong to no plane are copied unchanged into every level (``create_mixtures_from_indices`` :154-177), and
level d of the result is ``[unselected, plane 0 level d, plane 1 level d, ...]`` in that order.

Differences, all deliberate: arrays instead of nested lists; each plane's HEM is seeded (the reference's
libc generator just runs on from plane to plane, so its result depends on call history); the missing
``sh_degree`` argument of the reference's ``from_mixture`` call (:33) is not reproduced because no
GaussianModel is built here — callers get arrays (or ``registration_inputs`` via the context).
"""
from __future__ import annotations

import numpy as np

from .hem import HemContext, HemParams

_KEYS = ("xyz", "colors", "opacity", "cov6", "sh")


def merge_plane_inliers(xyz, colors, opacity, cov6, sh, plane_indices_list, params: HemParams | None = None,
                        ctx: HemContext | None = None, device: int = 0, seed: int = 0):
    """Returns a list of ``params.cluster_level`` dicts (xyz, colors, opacity, cov6, sh), finest first —
    the per-level arrays the reference hands to ``create_models_from_mixture``."""
    params = params or HemParams()
    arrays = dict(xyz=np.ascontiguousarray(xyz, np.float32), colors=np.ascontiguousarray(colors, np.float32),
                  opacity=np.ascontiguousarray(opacity, np.float32).reshape(-1),
                  cov6=np.ascontiguousarray(cov6, np.float32), sh=np.ascontiguousarray(sh, np.float32))
    n = arrays["xyz"].shape[0]
    planes = [np.asarray(ix, dtype=np.int64).reshape(-1) for ix in plane_indices_list]
    selected = np.concatenate(planes) if planes else np.zeros(0, np.int64)
    unselected = np.setdiff1d(np.arange(n, dtype=np.int64), selected)           # qt_plane_merging.py:93-100
    levels = [{k: [arrays[k][unselected]] for k in _KEYS} for _ in range(params.cluster_level)]
    own = ctx is None
    if own:
        ctx = HemContext(device, seed)
    try:
        for ix in planes:                                                       # process_all_planes :134-147
            if ix.size == 0:
                continue
            ctx.build(*(arrays[k][ix] for k in _KEYS), cluster_level=params.cluster_level,
                      hem_reduction=params.hem_reduction, distance_delta=params.distance_delta,
                      color_delta=params.color_delta, decay_rate=params.decay_rate)
            for depth in range(params.cluster_level):
                lv = ctx.level(depth + 1)
                for k in _KEYS:
                    levels[depth][k].append(lv[k])
    finally:
        if own:
            ctx.close()
    return [{k: np.concatenate(v, axis=0) for k, v in lvl.items()} for lvl in levels]
