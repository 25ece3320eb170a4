"""Drop-in stand-in for the reference's pybind11 module ``mixture_bind``
(reference src/cpp_ext/mixture_bind.cpp:11-60): same class names, same static
methods, same positional arguments, same per-level outputs — backed by the
B200 CUDA library through the C ABI (include/hem_b200.h).

    import gaussiansplattingregistration_b200.mixture_bind as mixture_bind
    level   = mixture_bind.MixtureLevel.CreateMixtureLevel(xyz, colors, opacities, covariance, features)
    pyramid = mixture_bind.MixtureCreator.CreateMixture(cluster_level, hem_reduction, distance_delta,
                                                        color_delta, decay_rate, level)
    xyz, colors, opacities, covariance, features = mixture_bind.MixtureLevel.CreatePythonLists(pyramid[0])

Differences from the reference, all deliberate:
* inputs may be nested lists (as the reference requires), numpy arrays or
  torch tensors; lists are converted in ONE numpy pass instead of per-element
  pybind casts (20 s per million splats in the reference);
* ``MixtureLevel`` keeps float32 arrays; ``pointSet`` etc. are array-backed,
  ``CreatePythonLists`` still returns nested Python lists, and
  ``CreateArrays`` returns the arrays without the list detour;
* nothing is leaked (the reference leaks its Mixture, mixture_wrapper.cpp:11);
* the parent-selection RNG is seeded (``MixtureCreator.seed``) instead of
  unseeded libc ``rand`` (base.hpp:44-56);
* empty input and ``clusterLevel == 0`` are defined (``[]``) instead of UB.
Wrong inner lengths raise ``RuntimeError`` with the reference's messages
(vec.hpp:92-94, 473-475).
"""
from __future__ import annotations

import threading

import numpy as np

from .hem import HemContext

__all__ = ["vec3", "smat3", "FeatureVector", "MixtureLevel", "MixtureCreator"]


class vec3:
    """reference include/vec.hpp:85-104, bound at mixture_bind.cpp:14-23"""
    __slots__ = ("x", "y", "z")

    def __init__(self, *a):
        if len(a) == 0:
            self.x = self.y = self.z = 0.0
        elif len(a) == 1:
            v = list(a[0])
            if len(v) != 3:
                raise RuntimeError("Python list must have exactly 3 elements.")
            self.x, self.y, self.z = (float(t) for t in v)
        elif len(a) == 3:
            self.x, self.y, self.z = (float(t) for t in a)
        else:
            raise TypeError("vec3(): incompatible constructor arguments")

    def __repr__(self):
        return "<vec3(%f, %f, %f)>" % (self.x, self.y, self.z)


class smat3:
    """reference include/vec.hpp:456-484, bound at mixture_bind.cpp:25-38"""
    __slots__ = ("e00", "e01", "e02", "e11", "e12", "e22")

    def __init__(self, *a):
        if len(a) == 0:
            v = [0.0] * 6
        elif len(a) == 1:
            v = list(a[0])
            if len(v) != 6:
                raise RuntimeError("Python list must have exactly 6 elements.")
        elif len(a) == 6:
            v = a
        else:
            raise TypeError("smat3(): incompatible constructor arguments")
        self.e00, self.e01, self.e02, self.e11, self.e12, self.e22 = (float(t) for t in v)

    def __repr__(self):
        return "<smat3(%f, %f, %f, %f, %f, %f)>" % (self.e00, self.e01, self.e02, self.e11, self.e12, self.e22)


class FeatureVector:
    """reference include/featurevector.hpp:8-26, bound at mixture_bind.cpp:40-43"""

    def __init__(self, a=None):
        if a is None:
            self._v = np.zeros(0, np.float32)
        elif isinstance(a, (int, np.integer)):
            self._v = np.zeros(int(a), np.float32)
        else:
            self._v = np.asarray(list(a), dtype=np.float32)

    def GetSize(self):
        return int(self._v.shape[0])

    def GetVector(self):
        return self._v.tolist()


def _as2d(a, width, what):
    """nested list / ndarray / tensor -> contiguous float32 [N, width]."""
    if type(a).__module__.startswith("torch"):
        a = a.detach().cpu().numpy()
    try:
        arr = np.asarray(a, dtype=np.float32)
    except ValueError as e:                     # ragged nested list
        if width in (3, 6):
            raise RuntimeError(f"Python list must have exactly {width} elements.") from e
        raise RuntimeError("feature vectors must all have the same length") from e
    if arr.ndim == 1 and arr.size == 0:
        arr = arr.reshape(0, width if width else 0)
    if arr.ndim != 2:
        raise RuntimeError(f"{what}: expected a list of N lists")
    if width and arr.shape[1] != width:
        raise RuntimeError(f"Python list must have exactly {width} elements.")
    return np.ascontiguousarray(arr)


class MixtureLevel:
    """reference include/mixturelevel.hpp:11-32, bound at mixture_bind.cpp:45-51"""

    def __init__(self):
        self.pointSet = np.zeros((0, 3), np.float32)
        self.colorSet = np.zeros((0, 3), np.float32)
        self.opacities = np.zeros(0, np.float32)
        self.covarianceSet = np.zeros((0, 6), np.float32)
        self.features = np.zeros((0, 0), np.float32)

    def __len__(self):
        return int(self.pointSet.shape[0])

    @staticmethod
    def CreateMixtureLevel(xyz, colors, opacities, covariance, features) -> "MixtureLevel":
        """reference src/cpp_ext/src/mixturelevel.cpp:14-28 — argument order xyz,
        colors, opacities, covariance, features."""
        m = MixtureLevel()
        m.pointSet = _as2d(xyz, 3, "xyz")
        m.colorSet = _as2d(colors, 3, "colors")
        if type(opacities).__module__.startswith("torch"):
            opacities = opacities.detach().cpu().numpy()
        m.opacities = np.ascontiguousarray(np.asarray(opacities, dtype=np.float32).reshape(-1))
        m.covarianceSet = _as2d(covariance, 6, "covariance")
        m.features = _as2d(features, 0, "features")
        n = m.pointSet.shape[0]
        if not (m.colorSet.shape[0] == m.opacities.shape[0] == m.covarianceSet.shape[0] == n):
            raise RuntimeError("xyz, colors, opacities and covariance must have the same length")
        if m.features.shape[0] != n:
            if m.features.size == 0:
                m.features = np.zeros((n, 0), np.float32)
            else:
                raise RuntimeError("features must have one row per Gaussian")
        return m

    @staticmethod
    def CreateArrays(mixtureLevel: "MixtureLevel"):
        """(xyz[N,3], colors[N,3], opacities[N], covariance[N,6], features[N,F]) float32 arrays."""
        return (mixtureLevel.pointSet, mixtureLevel.colorSet, mixtureLevel.opacities,
                mixtureLevel.covarianceSet, mixtureLevel.features)

    @staticmethod
    def CreatePythonLists(mixtureLevel: "MixtureLevel"):
        """reference src/cpp_ext/src/mixturelevel.cpp:30-70 — nested Python lists."""
        return tuple(a.tolist() for a in MixtureLevel.CreateArrays(mixtureLevel))


class MixtureCreator:
    """reference src/cpp_ext/mixture_wrapper.hpp:7-11, bound at mixture_bind.cpp:53-54"""

    #: seed of the counter-based parent-selection RNG (reference: unseeded libc rand)
    seed = 0
    #: CUDA device of the calling process
    device = 0
    _ctx = None
    _lock = threading.Lock()

    @classmethod
    def _context(cls) -> HemContext:
        if cls._ctx is None or cls._ctx.device != cls.device or cls._ctx_seed != cls.seed:
            if cls._ctx is not None:
                cls._ctx.close()
            cls._ctx = HemContext(cls.device, cls.seed)
            cls._ctx_seed = cls.seed
        return cls._ctx

    @classmethod
    def CreateMixture(cls, clusterLevel, hemReduction, distanceDelta, colorDelta, decayRate,
                      mixtureLevel: MixtureLevel):
        """reference src/cpp_ext/mixture_wrapper.cpp:10-18: returns ``clusterLevel``
        MixtureLevels, finest-reduced first (level 0, the input, is dropped)."""
        clusterLevel = int(clusterLevel)
        if clusterLevel < 0:
            raise TypeError("clusterLevel must be an unsigned integer")
        if clusterLevel == 0:
            return []
        with cls._lock:
            ctx = cls._context()
            ctx.build(mixtureLevel.pointSet, mixtureLevel.colorSet, mixtureLevel.opacities,
                      mixtureLevel.covarianceSet, mixtureLevel.features, clusterLevel,
                      float(hemReduction), float(distanceDelta), float(colorDelta), float(decayRate))
            out = []
            for l in range(1, clusterLevel + 1):
                d = ctx.level(l)
                m = MixtureLevel()
                m.pointSet, m.colorSet, m.opacities = d["xyz"], d["colors"], d["opacity"]
                m.covarianceSet, m.features = d["cov6"], d["sh"]
                out.append(m)
        return out
