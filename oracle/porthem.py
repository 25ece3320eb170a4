"""TEST INFRASTRUCTURE — ctypes access to oracle/libhem_oracle.so, the plain-C
CPU restatement (oracle/hem_oracle.c) of the reference HEM downsampler.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference
legs may import this module; the product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libhem_oracle.so")
_lib = None
_f = C.POINTER(C.c_float)
_u8 = C.POINTER(C.c_uint8)
_u32 = C.POINTER(C.c_uint32)
_i64 = C.POINTER(C.c_int64)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            build()
        L = C.CDLL(_PATH)
        L.hemo_run.restype = C.c_void_p
        L.hemo_run.argtypes = [_f, _f, _f, _f, _f, _f, _u8, C.c_int64, C.c_int, C.c_uint, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_uint64, C.c_uint64, C.c_int]
        L.hemo_num_levels.restype = C.c_int
        L.hemo_num_levels.argtypes = [C.c_void_p]
        L.hemo_level_size.restype = C.c_int64
        L.hemo_level_size.argtypes = [C.c_void_p, C.c_int]
        L.hemo_level_seconds.restype = C.c_double
        L.hemo_level_seconds.argtypes = [C.c_void_p, C.c_int]
        L.hemo_level_copy.restype = None
        L.hemo_level_copy.argtypes = [C.c_void_p, C.c_int, _f, _f, _f, _f, _f, _f, _u8]
        L.hemo_step_stats.restype = None
        L.hemo_step_stats.argtypes = [C.c_void_p, C.c_int, _i64]
        L.hemo_step_children.restype = None
        L.hemo_step_children.argtypes = [C.c_void_p, C.c_int, _u32, _i64, _u32, _f, _f, _f]
        L.hemo_free.restype = None
        L.hemo_free.argtypes = [C.c_void_p]
        L.hemo_philox_u32.restype = C.c_uint32
        L.hemo_philox_u32.argtypes = [C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def _p(a, t=_f):
    return None if a is None else a.ctypes.data_as(t)


def _c(a, dt=np.float32):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class PortRun:
    def __init__(self, h, F):
        self._h, self.F = h, F

    def close(self):
        if self._h:
            lib().hemo_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_levels(self):
        return lib().hemo_num_levels(self._h)

    def size(self, level):
        return int(lib().hemo_level_size(self._h, level))

    def seconds(self, step):
        return float(lib().hemo_level_seconds(self._h, step))

    def level(self, level) -> dict:
        n, F = self.size(level), self.F
        out = dict(xyz=np.zeros((n, 3), np.float32), colors=np.zeros((n, 3), np.float32),
                   opacity=np.zeros(n, np.float32), cov6=np.zeros((n, 6), np.float32),
                   sh=np.zeros((n, F), np.float32), weight=np.zeros(n, np.float32),
                   is_parent=np.zeros(n, np.uint8))
        lib().hemo_level_copy(self._h, level, _p(out["xyz"]), _p(out["colors"]), _p(out["opacity"]),
                              _p(out["cov6"]), _p(out["sh"]), _p(out["weight"]), _p(out["is_parent"], _u8))
        return out

    def stats(self, step) -> dict:
        s = np.zeros(5, np.int64)
        lib().hemo_step_stats(self._h, step, _p(s, _i64))
        return dict(nparents=int(s[0]), npairs=int(s[1]), norphans=int(s[2]), ndropped=int(s[3]),
                    nprefilter=int(s[4]))

    def children(self, step) -> dict:
        st = self.stats(step)
        n_in = self.size(step)
        out = dict(parent_idx=np.zeros(st["nparents"], np.uint32), off=np.zeros(st["nparents"] + 1, np.int64),
                   child=np.zeros(st["npairs"], np.uint32), wl=np.zeros(st["npairs"], np.float32),
                   sumlw=np.zeros(n_in, np.float32), radius=np.zeros(st["nparents"], np.float32))
        lib().hemo_step_children(self._h, step, _p(out["parent_idx"], _u32), _p(out["off"], _i64),
                                 _p(out["child"], _u32), _p(out["wl"]), _p(out["sumlw"]), _p(out["radius"]))
        return out


def run(xyz, colors, opacity, cov6, sh, levels, hem_reduction=3.0, distance_delta=3.0, color_delta=2.5,
        decay_rate=1.0, seed=0, weight=None, mask=None, rng_base=0, nthreads=0) -> PortRun:
    xyz, colors, opacity, cov6, sh = _c(xyz), _c(colors), _c(opacity), _c(cov6), _c(sh)
    weight, mask = _c(weight), _c(mask, np.uint8)
    n = xyz.shape[0]
    F = sh.shape[1] if sh.ndim == 2 else 0
    h = lib().hemo_run(_p(xyz), _p(colors), _p(opacity), _p(cov6), _p(sh), _p(weight), _p(mask, _u8), n, F,
                       levels, hem_reduction, distance_delta, color_delta, decay_rate, seed, rng_base, nthreads)
    return PortRun(h, F)


def philox_u32(seed: int, k: int) -> int:
    return int(lib().hemo_philox_u32(seed, k))
