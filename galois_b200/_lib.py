"""
ctypes binding of the C ABI declared in include/galois_b200.h.

The product path has NO CPU fallback: if libgalois_b200.so is missing or no CUDA device is
visible, the first arithmetic call raises (``BackendUnavailable``).  Nothing here imports
``oracle``.
"""

from __future__ import annotations

import ctypes
import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libgalois_b200.so"

GFB_OK, GFB_ERR_ZERO_DIVISION, GFB_ERR_BAD_ARG, GFB_ERR_CUDA, GFB_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
GFB_PRIME, GFB_BINARY = 1, 2
GFB_ADD, GFB_SUB, GFB_MUL, GFB_DIV = 0, 1, 2, 3
GFB_NEG, GFB_INV, GFB_SQUARE = 0, 1, 2

ZERO_DIVISION_MESSAGE = "Cannot compute the multiplicative inverse of 0 in a Galois field."


class BackendUnavailable(RuntimeError):
    """The CUDA extension is not built / not loadable, or there is no CUDA device."""


_vp, _sz, _u64, _i64, _int = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int64, C.c_int

# name -> (restype, argtypes); every symbol include/galois_b200.h declares
SIGNATURES = {
    "gfb_device_count": (_int, []),
    "gfb_last_error": (C.c_char_p, []),
    "gfb_version": (C.c_char_p, []),
    "gfb_take_error": (_int, [_vp, _vp]),
    "gfb_host_alloc": (_vp, [_sz]),
    "gfb_host_free": (None, [_vp]),
    "gfb_field_create": (_int, [C.POINTER(_vp), _int, _u64, _int, _u64, _u64, _u64]),
    "gfb_field_destroy": (None, [_vp]),
    "gfb_field_uses_tables": (_int, [_vp]),
    "gfb_binary": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz, _vp]),
    "gfb_binary_scalar": (_int, [_vp, _int, _int, _vp, _u64, _u64, _int, _vp, _sz, _vp]),
    "gfb_unary": (_int, [_vp, _int, _int, _vp, _vp, _sz, _vp]),
    "gfb_power": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _vp]),
    "gfb_power_scalar": (_int, [_vp, _int, _vp, _i64, _vp, _sz, _vp]),
    "gfb_binary_host": (_int, [_vp, _int, _int, _vp, _vp, _vp, _sz]),
    "gfb_binary_scalar_host": (_int, [_vp, _int, _int, _vp, _u64, _u64, _int, _vp, _sz]),
    "gfb_unary_host": (_int, [_vp, _int, _int, _vp, _vp, _sz]),
    "gfb_power_host": (_int, [_vp, _int, _vp, _vp, _vp, _sz]),
    "gfb_power_scalar_host": (_int, [_vp, _int, _vp, _i64, _vp, _sz]),
    "gfb_fold_workspace_bytes": (_sz, [_int, _sz, _sz, _sz]),
    "gfb_fold": (_int, [_vp, _int, _int, _int, _vp, _vp, _sz, _sz, _sz, _vp, _vp]),
    "gfb_fold_host": (_int, [_vp, _int, _int, _int, _vp, _vp, _sz, _sz, _sz]),
    "gfb_fold_segments": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gfb_fold_segments_host": (_int, [_vp, _int, _int, _vp, _sz, _vp, _vp, _vp, _sz]),
    "gfb_ntt_plan_create": (_int, [C.POINTER(_vp), _vp, _sz, _u64, _u64]),
    "gfb_ntt_plan_destroy": (None, [_vp]),
    "gfb_ntt_workspace_bytes": (_sz, [_vp, _int, _sz]),
    "gfb_ntt_exec": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _vp]),
    "gfb_ntt_exec_host": (_int, [_vp, _int, _vp, _vp, _sz]),
    "gfb_convolve": (_int, [_vp, _int, _vp, _sz, _vp, _sz, _vp, _vp]),
    "gfb_convolve_host": (_int, [_vp, _int, _vp, _sz, _vp, _sz, _vp]),
    "gfb_poly_evaluate": (_int, [_vp, _int, _vp, _sz, _vp, _vp, _sz, _vp]),
    "gfb_poly_evaluate_host": (_int, [_vp, _int, _vp, _sz, _vp, _vp, _sz]),
    "gfb_poly_roots_workspace_bytes": (_sz, [_vp, _sz]),
    "gfb_poly_roots": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    "gfb_poly_roots_host": (_int, [_vp, _int, _vp, _vp, _sz, _vp, _sz, C.POINTER(_sz)]),
    "gfb_matmul": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz, _sz, _sz, _int, _vp]),
    "gfb_matmul_host": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz, _sz, _sz, _int]),
    "gfb_row_reduce_workspace_bytes": (_sz, [_vp, _sz, _sz]),
    "gfb_row_reduce": (_int, [_vp, _int, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp]),
    "gfb_row_reduce_host": (_int, [_vp, _int, _vp, _vp, _sz, _sz, _sz, C.POINTER(_i64), _vp]),
    "gfb_plu_decompose": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _sz, _int, _vp, _vp, _vp]),
    "gfb_plu_decompose_host": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _sz, _sz, _int, C.POINTER(_i64), C.POINTER(_int)]),
    "gfb_rs_create": (_int, [C.POINTER(_vp), _vp, _int, _int, _int, _u64]),
    "gfb_bch_create": (_int, [C.POINTER(_vp), _vp, _int, _int, _int, _int, _u64, _vp]),
    "gfb_rs_destroy": (None, [_vp]),
    "gfb_rs_get_generator_poly": (_int, [_vp, _vp]),
    "gfb_rs_get_parity_matrix": (_int, [_vp, _vp]),
    "gfb_rs_encode": (_int, [_vp, _vp, _int, _vp, _sz, _vp]),
    "gfb_rs_decode": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _sz, _vp]),
    "gfb_rs_encode_host": (_int, [_vp, _vp, _int, _vp, _sz]),
    "gfb_rs_decode_host": (_int, [_vp, _vp, _vp, _int, _vp, _vp, _sz]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libgalois_b200.so (no CUDA call is made) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise BackendUnavailable(
            f"{LIB_PATH} is missing: build it with `python -m galois_b200.build` (needs nvcc). "
            "galois_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


_device_checked = False


def lib() -> ctypes.CDLL:
    """The loaded library, after verifying that a CUDA device is present (fail loudly otherwise)."""
    global _device_checked
    l = load()
    if not _device_checked:
        if l.gfb_device_count() <= 0:
            raise BackendUnavailable("no CUDA device is visible: galois_b200 has no CPU fallback")
        _device_checked = True
    return l


def last_error() -> str:
    return load().gfb_last_error().decode()


def check(status: int, what: str = "") -> None:
    """Map a gfb_status to the exception the reference raises for the same condition."""
    if status == GFB_OK:
        return
    if status == GFB_ERR_ZERO_DIVISION:
        raise ZeroDivisionError(ZERO_DIVISION_MESSAGE)
    if status == GFB_ERR_BAD_ARG:
        raise ValueError(f"galois_b200: invalid argument in {what}")
    if status == GFB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"galois_b200: {what} is not supported by the CUDA backend")
    raise RuntimeError(f"galois_b200: CUDA failure in {what}: {last_error()}")
