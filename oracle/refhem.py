"""TEST INFRASTRUCTURE — ctypes access to oracle/_ref/libhemref.so, i.e. the
UNMODIFIED reference cpp_ext compiled by oracle/Makefile plus the white-box
driver oracle/ref_driver.cpp.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libhemref.so")
_lib = None

_f = C.POINTER(C.c_float)
_u8 = C.POINTER(C.c_uint8)


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{_PATH} not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(_PATH, mode=C.RTLD_GLOBAL)
        L.refhem_run.restype = C.c_void_p
        L.refhem_run.argtypes = [_f, _f, _f, _f, _f, C.c_int64, C.c_int, C.c_uint, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_uint64, C.c_int]
        L.refhem_run_level.restype = C.c_void_p
        L.refhem_run_level.argtypes = [_f, _f, _f, _f, _f, _f, _u8, C.c_int64, C.c_int, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_int]
        L.refhem_num_levels.restype = C.c_int
        L.refhem_num_levels.argtypes = [C.c_void_p]
        L.refhem_level_size.restype = C.c_int64
        L.refhem_level_size.argtypes = [C.c_void_p, C.c_int]
        L.refhem_level_seconds.restype = C.c_double
        L.refhem_level_seconds.argtypes = [C.c_void_p, C.c_int]
        L.refhem_rng_calls.restype = C.c_uint64
        L.refhem_level_copy.restype = None
        L.refhem_level_copy.argtypes = [C.c_void_p, C.c_int, _f, _f, _f, _f, _f, _f, _u8]
        L.refhem_children.restype = C.c_int64
        L.refhem_children.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int64]
        L.refhem_radii.restype = None
        L.refhem_radii.argtypes = [C.c_void_p, C.c_int, C.c_float, _f]
        L.refhem_free.restype = None
        L.refhem_free.argtypes = [C.c_void_p]
        L.refhem_philox_u32.restype = C.c_uint32
        L.refhem_philox_u32.argtypes = [C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def _p(a, t=_f):
    return a.ctypes.data_as(t)


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


class RefRun:
    """Handle on one reference run (all levels retained, level 0 = input)."""

    def __init__(self, handle, F, dd):
        self._h = handle
        self.F = F
        self.dd = dd

    def close(self):
        if self._h:
            lib().refhem_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_levels(self) -> int:
        return lib().refhem_num_levels(self._h)

    def size(self, level: int) -> int:
        return int(lib().refhem_level_size(self._h, level))

    def seconds(self, clustering_level: int) -> float:
        return float(lib().refhem_level_seconds(self._h, clustering_level))

    def level(self, level: int) -> dict:
        n, F = self.size(level), self.F
        out = dict(xyz=np.zeros((n, 3), np.float32), colors=np.zeros((n, 3), np.float32),
                   opacity=np.zeros(n, np.float32), cov6=np.zeros((n, 6), np.float32),
                   sh=np.zeros((n, F), np.float32), weight=np.zeros(n, np.float32),
                   is_parent=np.zeros(n, np.uint8))
        lib().refhem_level_copy(self._h, level, _p(out["xyz"]), _p(out["colors"]), _p(out["opacity"]),
                                _p(out["cov6"]), _p(out["sh"]), _p(out["weight"]), _p(out["is_parent"], _u8))
        return out

    def children(self, level: int):
        """(parent_idx[Np], off[Np+1], child[P]) of the step consuming `level`."""
        L = lib()
        npar = C.c_int64(0)
        total = L.refhem_children(self._h, level, C.byref(npar), None, None, None, 0)
        off = np.zeros(npar.value + 1, np.int64)
        pidx = np.zeros(npar.value, np.uint32)
        child = np.zeros(max(total, 1), np.uint32)
        L.refhem_children(self._h, level, C.byref(npar), off.ctypes.data_as(C.POINTER(C.c_int64)),
                          pidx.ctypes.data_as(C.POINTER(C.c_uint32)),
                          child.ctypes.data_as(C.POINTER(C.c_uint32)), total)
        return pidx, off, child[:total]

    def radii(self, level: int) -> np.ndarray:
        out = np.zeros(self.size(level), np.float32)
        lib().refhem_radii(self._h, level, self.dd, _p(out))
        return out


def run(xyz, colors, opacity, cov6, sh, levels, hem_reduction=3.0, distance_delta=3.0, color_delta=2.5,
        decay_rate=1.0, seed=0, nthreads=0) -> RefRun:
    xyz, colors, opacity, cov6, sh = _c(xyz), _c(colors), _c(opacity), _c(cov6), _c(sh)
    n, F = xyz.shape[0], sh.shape[1]
    h = lib().refhem_run(_p(xyz), _p(colors), _p(opacity), _p(cov6), _p(sh), n, F, levels, hem_reduction,
                         distance_delta, color_delta, decay_rate, seed, nthreads)
    return RefRun(h, F, distance_delta)


def run_level(xyz, colors, opacity, cov6, sh, weight, is_parent, hem_reduction=3.0, distance_delta=3.0,
              color_delta=2.5, decay_rate=1.0, seed=0, rng_base=0, nthreads=0) -> RefRun:
    xyz, colors, opacity, cov6, sh = _c(xyz), _c(colors), _c(opacity), _c(cov6), _c(sh)
    weight = _c(weight)
    is_parent = _c(is_parent, np.uint8)
    n, F = xyz.shape[0], sh.shape[1]
    h = lib().refhem_run_level(_p(xyz), _p(colors), _p(opacity), _p(cov6), _p(sh), _p(weight),
                               _p(is_parent, _u8), n, F, hem_reduction, distance_delta, color_delta,
                               decay_rate, seed, rng_base, nthreads)
    return RefRun(h, F, distance_delta)


def philox_u32(seed: int, k: int) -> int:
    return int(lib().refhem_philox_u32(seed, k))
