"""
``np.convolve`` and ``np.matmul`` on field arrays, backed by gfb_convolve / gfb_matmul.

Mirrors ``convolve_jit.__call__`` (/root/reference/src/galois/_domains/_function.py:116-130) and
``matmul_jit.__call__`` (_domains/_linalg.py:190-276: 1-D promotion, batch broadcasting, squeezing).
"""

from __future__ import annotations

import numpy as np

from . import _lib, _pinned


def convolve(field, a, b, mode="full"):
    if not isinstance(a, field):
        raise TypeError(f"Argument 'a' must be a {field.name} array, not {type(a)}.")
    if not isinstance(b, field):
        raise TypeError(f"Argument 'b' must be a {field.name} array, not {type(b)}.")
    if mode != "full":
        raise ValueError(f"Operation 'convolve' currently only supports mode of 'full', not {mode!r}.")
    if a.ndim != 1 or b.ndim != 1:
        raise ValueError("Operation 'convolve' requires 1-D arrays.")
    if a.size == 0 or b.size == 0:
        raise ValueError("Operation 'convolve' requires non-empty arrays.")
    dtype = a.dtype
    av = field._to_dev_layout(field._prep(a, dtype))
    bv = field._to_dev_layout(field._prep(b, dtype))
    n = a.size + b.size - 1
    out = field._alloc_out(n, dtype)
    st = _lib.lib().gfb_convolve_host(field._handle(), field._elem_bytes(dtype), av.ctypes.data, a.size,
                                      bv.ctypes.data, b.size, out.ctypes.data)
    _lib.check(st, "gfb_convolve_host")
    return field._from_dev_layout(out, (n,), dtype)


def matmul(field, A, B):
    if not isinstance(A, field):
        raise TypeError(f"Argument 'A' must be a {field.name} array, not {type(A)}.")
    if not isinstance(B, field):
        raise TypeError(f"Argument 'B' must be a {field.name} array, not {type(B)}.")
    if not (A.ndim >= 1 and B.ndim >= 1):
        raise ValueError(f"Operation 'matmul' requires both arrays have dimension at least 1, not {A.ndim}-D and {B.ndim}-D.")
    dtype = A.dtype
    A_was_vec, B_was_vec = A.ndim == 1, B.ndim == 1
    Av = field._prep(A, dtype)
    Bv = field._prep(B, dtype)
    if A_was_vec:
        Av = Av.reshape(1, Av.size)
    if B_was_vec:
        Bv = Bv.reshape(Bv.size, 1)
    M, K = Av.shape[-2], Av.shape[-1]
    K2, N = Bv.shape[-2], Bv.shape[-1]
    if K != K2:
        raise ValueError(
            f"Operation 'matmul' requires the last dimension of 'A' to match the second-to-last dimension of 'B', "
            f"not {A.shape} and {B.shape}."
        )
    batch_shape = np.broadcast_shapes(Av.shape[:-2], Bv.shape[:-2])
    batch = int(np.prod(batch_shape)) if batch_shape else 1
    b_batched = 1 if Bv.ndim > 2 and int(np.prod(Bv.shape[:-2])) > 1 else 0
    A_flat = np.broadcast_to(Av, batch_shape + (M, K)).reshape(batch, M, K)
    B_flat = np.broadcast_to(Bv, batch_shape + (K, N)).reshape(batch, K, N) if b_batched else Bv.reshape(K, N)
    a_dev = field._to_dev_layout(A_flat)
    b_dev = field._to_dev_layout(B_flat)
    out = field._alloc_out(batch * M * N, dtype)
    if batch * M * N:
        st = _lib.lib().gfb_matmul_host(field._handle(), field._elem_bytes(dtype), a_dev.ctypes.data, b_dev.ctypes.data,
                                        out.ctypes.data, batch, M, K, N, b_batched)
        _lib.check(st, "gfb_matmul_host")
    if A_was_vec and B_was_vec:
        final = batch_shape
    elif A_was_vec:
        final = batch_shape + (N,)
    elif B_was_vec:
        final = batch_shape + (M,)
    else:
        final = batch_shape + (M, N)
    C = field._from_dev_layout(out, final, dtype)
    if C.ndim == 0:
        return field(int(C.item()), dtype=dtype)
    return C


def poly_evaluate(coeffs, x):
    """
    Evaluate the polynomial with degree-descending coefficients ``coeffs`` at every element of ``x`` (Horner).
    Mirrors ``evaluate_elementwise_jit.__call__`` (/root/reference/src/galois/_polys/_dense.py:409-424), the kernel
    behind ``Poly.__call__``: result has ``x``'s shape and dtype.
    """
    from ._field import FieldArray

    if not isinstance(coeffs, FieldArray) or not isinstance(x, FieldArray) or type(coeffs) is not type(x):
        raise TypeError("Arguments 'coeffs' and 'x' must be arrays over the same field.")
    field = type(x)
    if coeffs.ndim != 1 or coeffs.size == 0:
        raise ValueError("Argument 'coeffs' must be a non-empty 1-D array.")
    dtype = x.dtype
    shape = x.shape
    xv = field._to_dev_layout(field._prep(np.atleast_1d(x).reshape(-1).view(field), dtype))
    cv = field._to_dev_layout(field._prep(coeffs, dtype))
    n = int(np.prod(shape)) if shape else 1
    out = field._alloc_out(n, dtype)
    if n:
        st = _lib.lib().gfb_poly_evaluate_host(field._handle(), field._elem_bytes(dtype), cv.ctypes.data, coeffs.size,
                                               xv.ctypes.data, out.ctypes.data, n)
        _lib.check(st, "gfb_poly_evaluate_host")
    res = field._from_dev_layout(out, shape, dtype)
    return res if res.ndim else field(int(res.item()), dtype=dtype)
