"""ctypes binding of libhem_b200.so (C ABI declared in include/hem_b200.h).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``)
and must be present: there is no CPU fallback, a missing or unloadable
extension raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEM_B200_LIB") or os.path.join(_HERE, "libhem_b200.so")   # override: kernel experiments only

STEP_STATS_LEN = 24
ICP_RESULT_LEN = 24
NCCL_ID_BYTES = 128

_f = C.POINTER(C.c_float)
_u8 = C.POINTER(C.c_uint8)
_u32 = C.POINTER(C.c_uint32)
_i64 = C.POINTER(C.c_int64)
_d = C.POINTER(C.c_double)

#: every symbol include/hem_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "hem_version": (C.c_int, []),
    "hem_create": (C.c_void_p, [C.c_int, C.c_uint64]),
    "hem_destroy": (None, [C.c_void_p]),
    "hem_last_error": (C.c_char_p, [C.c_void_p]),
    "hem_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hem_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "hem_build_pyramid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_void_p]),
    "hem_num_levels": (C.c_int, [C.c_void_p]),
    "hem_level_size": (C.c_int64, [C.c_void_p, C.c_int]),
    "hem_level_copy": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "hem_level_registration_inputs": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hem_level_from_mixture": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hem_icp_point_to_point": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_double,
                                         C.c_double, C.c_double, C.c_int, _d, C.c_void_p]),
    "hem_icp_multiscale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _d, C.POINTER(C.c_int32), C.c_double, C.c_double,
                                     C.c_void_p, _d]),
    "hem_bind_output": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "hem_clear_outputs": (C.c_int, [C.c_void_p]),
    "hem_step_stats": (C.c_int, [C.c_void_p, C.c_int, _d]),
    "hem_get_info": (C.c_int, [C.c_void_p, C.c_char_p, _d]),
    "hem_step_pairs": (C.c_int, [C.c_void_p, C.c_int, _u32, _i64, _u32, _f, _f, _f]),
    "hem_nccl_unique_id": (C.c_int, [_u8]),
    "hem_comm_init": (C.c_int, [C.c_void_p, _u8, C.c_int, C.c_int]),
    "hem_slab_range": (C.c_int, [C.c_int64, C.c_int, C.c_int, _i64, _i64]),
}

_lib = None


class HemError(RuntimeError):
    """An error reported by libhem_b200 (code + hem_last_error message)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libhem_b200 error {code}: {message}")
        self.code = code


def load():
    """Load the shared library and bind every declared symbol (no compute)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
