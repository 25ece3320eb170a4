"""Headless mirror of the reference's local-registration entry points, on the GPU.

SURVEY.md §8f rank 4.  The reference registers pyramid levels with Open3D (third-party, pinned
0.16.0, not vendored) through two functions:

* ``do_icp_registration(point_cloud_first, point_cloud_second, init_transform, registration_params)``
  (src/utils/local_registration_util.py:76-100), parameters ``LocalRegistrationParams``
  (src/params/registration_parameters.py:8-15);
* the coarse-to-fine loop ``MultiScaleRegistratorMixture._register_main_point_clouds``
  (src/gui/workers/registration/qt_multiscale_registrator.py:203-230), which calls the former with ten
  positional arguments although it takes four — ``multiscale_registration`` below
  is that loop with the call repaired.

Same names, argument order and result fields (``transformation``, ``fitness``, ``inlier_rmse``,
``correspondence_set``).  Only the reference's default estimator — point-to-point, no robust loss —
is on this path; the other registration types need normals/colour terms and raise NotImplementedError
(they are not silently replaced).  There is no CPU path: everything runs through ``libhem_b200.so``.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum

import numpy as np

from .hem import HemContext, IcpResult


class LocalRegistrationType(Enum):
    """src/utils/local_registration_util.py:22-34 (values = the reference's instance names)."""
    ICP_Point_To_Point = "Point-to-Point ICP"
    ICP_Point_To_Plane = "Point-to-Plane ICP"
    ICP_Color = "Colored ICP"
    ICP_General = "Generalized ICP"

    @property
    def instance_name(self) -> str:
        return self.value


class KernelLossFunctionType(Enum):
    """src/utils/local_registration_util.py:6-19."""
    Loss_None = "None"
    Tukey_Loss = "Tukey loss"
    Cauchy_Loss = "Cauchy loss"
    GMLoss = "GM loss"
    Huber_Loss = "Huber loss"

    @property
    def instance_name(self) -> str:
        return self.value


@dataclass
class LocalRegistrationParams:
    """src/params/registration_parameters.py:8-15, same defaults."""
    registration_type: LocalRegistrationType = LocalRegistrationType.ICP_Point_To_Point
    max_correspondence: float = 5.0
    relative_fitness: float = 0.000001
    relative_rmse: float = 0.000001
    max_iteration: int = 30
    rejection_type: KernelLossFunctionType = KernelLossFunctionType.Loss_None
    k_value: float = 0.0


def _points_of(pc):
    """An Open3D-like point cloud (``.points``), a dict from ``HemContext.registration_inputs`` or an [n,3] array."""
    if hasattr(pc, "points"):
        pc = pc.points
    elif isinstance(pc, dict):
        pc = pc["points"]
    if type(pc).__module__.startswith("torch"):
        return pc
    return np.ascontiguousarray(np.asarray(pc, np.float64).reshape(-1, 3))


def do_icp_registration(point_cloud_first, point_cloud_second, init_transform,
                        registration_params: LocalRegistrationParams, ctx: HemContext | None = None) -> IcpResult:
    """Drop-in for the reference's function of the same name (first = source, second = target)."""
    p = registration_params
    if p.registration_type is not LocalRegistrationType.ICP_Point_To_Point:
        raise NotImplementedError(f"{p.registration_type.instance_name} is not on the B200 path "
                                  "(only the reference's default, point-to-point ICP, is)")
    own = ctx is None
    if own:
        ctx = HemContext(0, 0)
    try:
        return ctx.icp(_points_of(point_cloud_first), _points_of(point_cloud_second), init_transform,
                       p.max_correspondence, p.relative_fitness, p.relative_rmse, p.max_iteration,
                       with_correspondences=True)
    finally:
        if own:
            ctx.close()


def multiscale_registration(pc1_list, pc2_list, init_trans, voxel_values, iter_values,
                            relative_fitness: float = 1e-6, relative_rmse: float = 1e-6,
                            registration_type: LocalRegistrationType = LocalRegistrationType.ICP_Point_To_Point):
    """The reference's coarse-to-fine loop with its validity checks (qt_multiscale_registrator.py:170-230).

    ``pc1_list`` / ``pc2_list`` are either two ``HemContext`` objects holding built pyramids (then nothing but the
    results leaves the GPU) or two lists ``[full cloud, level 1, ...]`` of point clouds / arrays.
    Returns the list of per-scale results, coarsest first; the last one is the reference's return value.
    """
    if registration_type is not LocalRegistrationType.ICP_Point_To_Point:
        raise NotImplementedError(f"{registration_type.instance_name} is not on the B200 path")
    if len(iter_values) != len(voxel_values):
        raise ValueError("The number of iteration and voxel values provided do not match.")
    if isinstance(pc1_list, HemContext) and isinstance(pc2_list, HemContext):
        return pc1_list.icp_multiscale(pc2_list, voxel_values, iter_values, init_trans, relative_fitness, relative_rmse)
    if len(pc1_list) != len(pc2_list):
        raise ValueError("The two point cloud lists differ in size.")
    if len(pc1_list) <= 1:
        raise ValueError("There are no downscaled mixtures. First create Gaussian Mixtures by running the HEM "
                         "algorithm in the \"Mixture\" tab!")
    if len(iter_values) != len(pc1_list):
        raise ValueError("The number of iterations and the mixture levels do not match.")
    out = []
    current = np.eye(4) if init_trans is None else np.asarray(init_trans, np.float64)
    with HemContext(0, 0) as ctx:
        for index in range(len(iter_values)):
            params = LocalRegistrationParams(max_correspondence=voxel_values[index], relative_fitness=relative_fitness,
                                             relative_rmse=relative_rmse, max_iteration=iter_values[index])
            r = do_icp_registration(pc1_list[-(index + 1)], pc2_list[-(index + 1)], current, params, ctx)
            current = r.transformation
            out.append(r)
    return out


def register_gaussian_clouds(first, second, hem_params=None, voxel_values=(5.0, 2.5, 2.0, 1.0), iter_values=(50, 30, 20, 10),
                             init_trans=None, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6,
                             seed: int = 0, device: int = 0):
    """BASELINE configs[4] as one headless call: what the GUI does with its "Mixture" tab (HEM pyramid of each cloud,
    qt_gaussian_mixture.py:42-58) followed by its "Multiscale" tab (coarse-to-fine ICP over the two pyramids,
    qt_multiscale_registrator.py:203-230), everything staying on the GPU in between.

    ``first`` / ``second``: tuples ``(xyz, colors, opacity, cov6, sh)`` of numpy arrays or torch CUDA tensors.
    ``len(voxel_values) == len(iter_values) == hem_params.cluster_level + 1`` as in the reference.
    The two pyramids are independent, so they are built from two host threads on the two contexts' streams (the C ABI
    drops the GIL): their launch-bound coarse levels overlap on the one GPU.
    Returns ``(results, ctx_first, ctx_second)``; the caller closes the contexts (they hold the pyramids).
    """
    import threading
    from .hem import HemParams
    p = hem_params or HemParams()
    if len(iter_values) != len(voxel_values):
        raise ValueError("The number of iteration and voxel values provided do not match.")
    if len(iter_values) != p.cluster_level + 1:
        raise ValueError("The number of iterations and the mixture levels do not match.")
    ctxs = (HemContext(device, seed), HemContext(device, seed))
    errs = []

    def build(ctx, cloud):
        try:
            ctx.build(*cloud, cluster_level=p.cluster_level, hem_reduction=p.hem_reduction, distance_delta=p.distance_delta,
                      color_delta=p.color_delta, decay_rate=p.decay_rate)
        except Exception as e:          # surfaced on the calling thread below
            errs.append(e)
    threads = [threading.Thread(target=build, args=(c, x)) for c, x in zip(ctxs, (first, second))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        for c in ctxs:
            c.close()
        raise errs[0]
    results = ctxs[0].icp_multiscale(ctxs[1], list(voxel_values), list(iter_values), init_trans, relative_fitness, relative_rmse)
    return results, ctxs[0], ctxs[1]
