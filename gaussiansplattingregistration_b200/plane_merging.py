"""Second caller of the HEM boundary: plane-inlier merging.

Mirror of the array logic of the reference's ``PlaneInlierMergingWorker``
(src/gui/workers/downsampling/qt_plane_merging.py:86-177) without the Qt shell: the inliers of every
fitted plane are reduced by their own HEM pyramid (``process_all_planes`` :132-152), the splats that
belong to no plane are copied unchanged into every level (``create_mixtures_from_indices`` :154-177), and
level d of the result is ``[unselected, plane 0 level d, plane 1 level d, ...]`` in that order.

The planes are independent problems, each far too small to fill a B200 (a few thousand to a few hundred thousand
inliers: a launch-bound pyramid of ~20 dependent kernels per level): instead of one pyramid after the other they are
built CONCURRENTLY, one context (own stream, own workspace) per worker thread, the C ABI call releasing the GIL — the
GPU interleaves the planes' kernels, and the result order stays the reference's.

Differences, all deliberate: arrays instead of nested lists; each plane's HEM is seeded (the reference's
libc generator just runs on from plane to plane, so its result depends on call history); the missing
``sh_degree`` argument of the reference's ``from_mixture`` call (:33) is not reproduced because no
GaussianModel is built here — callers get arrays (or ``registration_inputs`` via the context).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .hem import HemContext, HemParams

_KEYS = ("xyz", "colors", "opacity", "cov6", "sh")


def merge_plane_inliers(xyz, colors, opacity, cov6, sh, plane_indices_list, params: HemParams | None = None,
                        ctx: HemContext | None = None, device: int = 0, seed: int = 0, concurrency: int = 4):
    """Returns a list of ``params.cluster_level`` dicts (xyz, colors, opacity, cov6, sh), finest first —
    the per-level arrays the reference hands to ``create_models_from_mixture``.  With ``ctx`` the planes are built one
    after the other on that context; otherwise up to ``concurrency`` contexts build them side by side (same seed for
    every plane either way, so both give the same arrays)."""
    params = params or HemParams()
    arrays = dict(xyz=np.ascontiguousarray(xyz, np.float32), colors=np.ascontiguousarray(colors, np.float32),
                  opacity=np.ascontiguousarray(opacity, np.float32).reshape(-1),
                  cov6=np.ascontiguousarray(cov6, np.float32), sh=np.ascontiguousarray(sh, np.float32))
    n = arrays["xyz"].shape[0]
    planes = [np.asarray(ix, dtype=np.int64).reshape(-1) for ix in plane_indices_list]
    selected = np.concatenate(planes) if planes else np.zeros(0, np.int64)
    unselected = np.setdiff1d(np.arange(n, dtype=np.int64), selected)           # qt_plane_merging.py:93-100
    levels = [{k: [arrays[k][unselected]] for k in _KEYS} for _ in range(params.cluster_level)]
    work = [i for i, ix in enumerate(planes) if ix.size]                        # process_all_planes :134-147 skips nothing else

    def build_plane(c: HemContext, i: int):
        ix = planes[i]
        c.build(*(arrays[k][ix] for k in _KEYS), cluster_level=params.cluster_level,
                hem_reduction=params.hem_reduction, distance_delta=params.distance_delta,
                color_delta=params.color_delta, decay_rate=params.decay_rate)
        return [c.level(depth + 1) for depth in range(params.cluster_level)]

    results = {}
    if ctx is not None or len(work) <= 1 or concurrency <= 1:
        own = ctx is None
        c = ctx if ctx is not None else HemContext(device, seed)
        try:
            for i in work:
                results[i] = build_plane(c, i)
        finally:
            if own:
                c.close()
    else:
        pool = [HemContext(device, seed) for _ in range(min(concurrency, len(work)))]
        try:
            def worker(slot: int):
                out = {}
                for i in work[slot::len(pool)]:                                 # static round-robin: deterministic and balanced enough
                    out[i] = build_plane(pool[slot], i)
                return out
            with ThreadPoolExecutor(max_workers=len(pool)) as ex:
                for part in ex.map(worker, range(len(pool))):
                    results.update(part)
        finally:
            for c in pool:
                c.close()
    for i in work:                                                              # reference order: plane 0, plane 1, ...
        for depth in range(params.cluster_level):
            for k in _KEYS:
                levels[depth][k].append(results[i][depth][k])
    return [{k: np.concatenate(v, axis=0) for k, v in lvl.items()} for lvl in levels]
