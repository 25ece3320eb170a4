"""Seeded synthetic 3D-Gaussian-splat clouds for parity tests and benchmarks.

The reference ships no sample data, so every test and bench in
this repository uses this generator (SURVEY.md §8d "Configs -> concrete
inputs").  Output layout is exactly what the reference's caller hands to
``mixture_bind.MixtureLevel.CreateMixtureLevel`` (reference
src/gui/workers/downsampling/qt_gaussian_mixture.py:42-47):

* ``xyz``      [N,3]  means
* ``colors``   [N,3]  raw SH-DC coefficients
* ``opacity``  [N]    raw (pre-sigmoid) opacity logits
* ``cov6``     [N,6]  packed covariance, order 00,01,02,11,12,22
                       (reference src/utils/general_utils.py:20-29)
* ``sh``       [N,F]  SH "rest" coefficients, F = 3*((deg+1)^2 - 1)

Pure numpy, float32, ``numpy.random.default_rng(seed)``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SplatCloud:
    xyz: np.ndarray
    colors: np.ndarray
    opacity: np.ndarray
    cov6: np.ndarray
    sh: np.ndarray

    @property
    def n(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def sh_floats(self) -> int:
        return int(self.sh.shape[1])

    def arrays(self):
        return self.xyz, self.colors, self.opacity, self.cov6, self.sh


def sh_rest_floats(sh_degree: int) -> int:
    return 3 * ((sh_degree + 1) ** 2 - 1)


def synth(n: int, seed: int = 0, density: float = 5000.0, sh_degree: int = 3,
          aspect=(1.0, 1.0, 1.0), scale_median: float = 0.03, scale_sigma_ln: float = 0.5,
          clip_sigma: float | None = 3.0) -> SplatCloud:
    """Uniform positions in a box of volume n/density with the given aspect,
    log-normal scales (clipped at +-clip_sigma*sigma_ln unless None), random
    rotations, N(0,1) colours, N(0,2) opacity logits, N(0,0.1) SH rest."""
    rng = np.random.default_rng(seed)
    n = int(n)
    vol = n / float(density) if n > 0 else 1.0
    a = np.asarray(aspect, dtype=np.float64)
    unit = (vol / a.prod()) ** (1.0 / 3.0)
    ext = (a * unit).astype(np.float32)
    xyz = (rng.random((n, 3), dtype=np.float32) * ext).astype(np.float32)

    z = rng.standard_normal((n, 3)).astype(np.float32)
    if clip_sigma is not None:
        z = np.clip(z, -clip_sigma, clip_sigma)
    s = (scale_median * np.exp(scale_sigma_ln * z)).astype(np.float32)

    q = rng.standard_normal((n, 4)).astype(np.float64)
    q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)
    w, x, y, zq = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 3, 3), dtype=np.float64)
    R[:, 0, 0] = 1 - 2 * (y * y + zq * zq)
    R[:, 0, 1] = 2 * (x * y - w * zq)
    R[:, 0, 2] = 2 * (x * zq + w * y)
    R[:, 1, 0] = 2 * (x * y + w * zq)
    R[:, 1, 1] = 1 - 2 * (x * x + zq * zq)
    R[:, 1, 2] = 2 * (y * zq - w * x)
    R[:, 2, 0] = 2 * (x * zq - w * y)
    R[:, 2, 1] = 2 * (y * zq + w * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    M = R * s[:, None, :].astype(np.float64)          # R * diag(s)
    C = M @ np.transpose(M, (0, 2, 1))                # R diag(s^2) R^T
    cov6 = np.stack([C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2]],
                    axis=1).astype(np.float32)

    colors = rng.standard_normal((n, 3)).astype(np.float32)
    opacity = (2.0 * rng.standard_normal(n)).astype(np.float32)
    F = sh_rest_floats(sh_degree)
    sh = (0.1 * rng.standard_normal((n, F))).astype(np.float32)
    return SplatCloud(np.ascontiguousarray(xyz), np.ascontiguousarray(colors),
                      np.ascontiguousarray(opacity), np.ascontiguousarray(cov6),
                      np.ascontiguousarray(sh))


def rigid_copy(cloud: SplatCloud, seed: int = 1, max_rot_deg: float = 10.0,
               max_trans_frac: float = 0.05, noise: float = 0.01):
    """Second cloud of BASELINE config 5: the same splats under a known rigid
    transform with independent SH/opacity noise.  Returns (cloud2, T[4,4])."""
    rng = np.random.default_rng(seed)
    axis = rng.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(max_rot_deg) * rng.random()
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    ext = cloud.xyz.max(0) - cloud.xyz.min(0) if cloud.n else np.ones(3)
    t = (rng.random(3) * 2 - 1) * max_trans_frac * ext
    xyz = (cloud.xyz.astype(np.float64) @ R.T + t).astype(np.float32)
    c = cloud.cov6.astype(np.float64)
    C = np.empty((cloud.n, 3, 3))
    C[:, 0, 0], C[:, 0, 1], C[:, 0, 2] = c[:, 0], c[:, 1], c[:, 2]
    C[:, 1, 0], C[:, 1, 1], C[:, 1, 2] = c[:, 1], c[:, 3], c[:, 4]
    C[:, 2, 0], C[:, 2, 1], C[:, 2, 2] = c[:, 2], c[:, 4], c[:, 5]
    C = R @ C @ R.T
    cov6 = np.stack([C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2]],
                    axis=1).astype(np.float32)
    sh = (cloud.sh + noise * rng.standard_normal(cloud.sh.shape)).astype(np.float32)
    op = (cloud.opacity + noise * rng.standard_normal(cloud.n)).astype(np.float32)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return SplatCloud(xyz, cloud.colors.copy(), op, cov6, sh), T
