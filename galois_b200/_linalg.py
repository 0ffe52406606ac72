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


# ------------------------------------------------------------------------------------------------
# Gaussian elimination: row_reduce / lu / plu and what the reference builds from them
# ------------------------------------------------------------------------------------------------
def _no_wide(field, what):
    if field._wide:
        raise NotImplementedError(f"{what} is not implemented over {field.name} (elements wider than 64 bits).")


def _row_reduce_raw(field, A, ncols=None, want_det=False):
    """gfb_row_reduce_host on a 2-D field array -> (A_rre, n_pivots[, det]); row_reduce_jit.__call__ (_linalg.py:321-353)."""
    import ctypes

    if not isinstance(A, field):
        raise TypeError(f"Argument 'A' must be a {field.name} array, not {type(A)}.")
    if not A.ndim == 2:
        raise ValueError(f"Only 2-D matrices can be converted to reduced row echelon form, not {A.ndim}-D.")
    _no_wide(field, "row_reduce")
    rows, cols = A.shape
    ncols = cols if ncols is None else int(ncols)
    if not 0 <= ncols <= cols:
        raise ValueError(f"Argument 'ncols' must be in [0, {cols}], not {ncols}.")
    dtype = A.dtype
    src = np.ascontiguousarray(field._prep(A, dtype))
    out = _pinned.empty((rows, cols), dtype)
    n_piv = ctypes.c_int64(0)
    det = np.zeros(1, dtype=dtype)
    st = _lib.lib().gfb_row_reduce_host(field._handle(), field._elem_bytes(dtype), src.ctypes.data, out.ctypes.data, rows,
                                        cols, ncols, ctypes.byref(n_piv), det.ctypes.data if want_det else None)
    _lib.check(st, "gfb_row_reduce_host")
    R = np.array(out, dtype=dtype).view(field)
    if want_det:
        return R, int(n_piv.value), det
    return R, int(n_piv.value)


def row_reduce(A, ncols=None, eye="left"):
    """FieldArray.row_reduce (_fields/_array.py:1412-1469)."""
    if eye not in ("left", "right"):
        raise ValueError(f"Argument 'eye' must be in ['left', 'right'], not {eye!r}.")
    field = type(A)
    if eye == "left":
        return _row_reduce_raw(field, A, ncols)[0]
    R = _row_reduce_raw(field, A[::-1, ::-1] if A.ndim == 2 else A, ncols)[0]
    return R[::-1, ::-1]


def _plu_raw(field, A, pivoting: bool):
    import ctypes

    if not isinstance(A, field):
        raise TypeError(f"Argument 'A' must be a {field.name} array, not {type(A)}.")
    if not A.ndim == 2:
        raise ValueError(f"Argument 'A' must be a 2-D matrix, not have shape {A.shape}.")
    _no_wide(field, "LU decomposition")
    m, n = A.shape
    if m == 0 or n == 0:
        raise ValueError(f"Argument 'A' must be a non-empty matrix, not have shape {A.shape}.")
    if not pivoting and m - 1 > n:
        # lu_decompose_jit walks the diagonal for range(m - 1) (_linalg.py:371-372): Ai[i, i] leaves the matrix
        raise IndexError(f"index {n} is out of bounds for axis 1 with size {n}")
    dtype = A.dtype
    src = np.ascontiguousarray(field._prep(A, dtype))
    U = _pinned.empty((m, n), dtype)
    L = _pinned.empty((m, m), dtype)
    perm = np.zeros(m, dtype=np.int64)
    n_swaps = ctypes.c_int64(0)
    failed = ctypes.c_int(0)
    st = _lib.lib().gfb_plu_decompose_host(field._handle(), field._elem_bytes(dtype), src.ctypes.data, U.ctypes.data,
                                           L.ctypes.data, perm.ctypes.data, m, n, 1 if pivoting else 0,
                                           ctypes.byref(n_swaps), ctypes.byref(failed))
    _lib.check(st, "gfb_plu_decompose_host")
    if failed.value:
        raise ValueError("The LU decomposition of 'A' does not exist. Use the PLU decomposition instead.")
    return (np.array(L, dtype=dtype).view(field), np.array(U, dtype=dtype).view(field), perm, int(n_swaps.value))


def lu_decompose(A):
    """FieldArray.lu_decompose (_fields/_array.py:1471-1502) -> (L, U)."""
    L, U, _, _ = _plu_raw(type(A), A, pivoting=False)
    return L, U


def plu_decompose(A):
    """FieldArray.plu_decompose (_fields/_array.py:1504-1539) -> (P, L, U) with A = P L U."""
    field = type(A)
    L, U, perm, _ = _plu_raw(field, A, pivoting=True)
    m = A.shape[0]
    P_rows = np.zeros((m, m), dtype=A.dtype)
    P_rows[np.arange(m), perm] = 1  # the reference's row-swapped identity; it returns the transpose (:427-428)
    return P_rows.T.copy().view(field), L, U


def _require_square(A):
    if not (A.ndim == 2 and A.shape[0] == A.shape[1]):
        raise np.linalg.LinAlgError(f"Argument 'A' must be square, not {A.shape}.")


def matrix_rank(field, A):
    """matrix_rank_jit (_linalg.py:479-489): non-zero rows of the reduced row echelon form."""
    R, _ = _row_reduce_raw(field, A)
    return int(np.sum(~np.all(np.asarray(R) == 0, axis=1)))


def inv(field, A):
    """inv_jit (_linalg.py:492-522): row-reduce [A | I] over the first n columns."""
    if not isinstance(A, field):
        raise TypeError(f"Argument 'A' must be a {field.name} array, not {type(A)}.")
    _require_square(A)
    n = A.shape[0]
    AI = np.concatenate((np.asarray(A), np.identity(n, dtype=A.dtype)), axis=-1).view(field)
    R, _ = _row_reduce_raw(field, AI, ncols=n)
    rank = int(np.sum(~np.all(np.asarray(R)[:, 0:n] == 0, axis=1)))
    if not rank == n:
        raise np.linalg.LinAlgError(
            f"Argument 'A' is singular and not invertible because it does not have full rank of {n}, but rank of {rank}."
        )
    return R[:, -n:] if n else R[:, 0:0]


def solve(field, A, b):
    """solve_jit (_linalg.py:525-546): x = A^-1 b."""
    if not isinstance(A, field):
        raise TypeError(f"Argument 'A' must be a {field.name} array, not {type(A)}.")
    if not isinstance(b, field):
        raise TypeError(f"Argument 'b' must be a {field.name} array, not {type(b)}.")
    _require_square(A)
    if b.ndim not in [1, 2]:
        raise np.linalg.LinAlgError(f"Argument 'b' must have dimension equal to 'A' or one less, not {b.ndim}.")
    if not A.shape[-1] == b.shape[0]:
        raise np.linalg.LinAlgError(
            f"The last dimension of 'A' must equal the first dimension of 'b', not {A.shape} and {b.shape}."
        )
    A_inv = inv(field, A)
    return matmul(field, np.ascontiguousarray(A_inv).view(field), b)


def det(field, A):
    """
    det_jit (_linalg.py:446-471).  The reference takes closed forms for n = 2, 3 and (-1)^swaps det(L) det(U) otherwise;
    the determinant is one field element whichever way it is computed — here it is the signed product of the pivots
    of the elimination (0 when the matrix is rank deficient).
    """
    if not isinstance(A, field):
        raise TypeError(f"Argument 'A' must be a {field.name} array, not {type(A)}.")
    _require_square(A)
    n = A.shape[0]
    if n == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")  # plu_decompose_jit's L[-1, -1] (:423)
    _, p, d = _row_reduce_raw(field, A, want_det=True)
    return field(int(d[0]) if p == n else 0, dtype=A.dtype)


def row_space(A):
    """FieldArray.row_space (_fields/_array.py:1541-1590)."""
    field = type(A)
    if not A.ndim == 2:
        raise ValueError(f"Only 2-D matrices have a row space, not {A.ndim}-D.")
    R = row_reduce(A)
    rank = int(np.sum(~np.all(np.asarray(R) == 0, axis=1)))
    return R[0:rank, :].view(field)


def column_space(A):
    """FieldArray.column_space (_fields/_array.py:1592-1637)."""
    if not A.ndim == 2:
        raise ValueError(f"Only 2-D matrices have a column space, not {A.ndim}-D.")
    return row_space(A.T)


def left_null_space(A):
    """FieldArray.left_null_space (_fields/_array.py:1639-1703): the rows of [A | I] that reduce to zero on the left."""
    field = type(A)
    if not A.ndim == 2:
        raise ValueError(f"Only 2-D matrices have a left null space, not {A.ndim}-D.")
    m, n = A.shape
    AI = np.concatenate((np.asarray(A), np.identity(m, dtype=A.dtype)), axis=-1).view(field)
    R, p = _row_reduce_raw(field, AI, ncols=n)
    LN = np.ascontiguousarray(np.asarray(R)[p:, n:]).view(field)
    if LN.shape[0] == 0:
        return LN
    return row_reduce(LN)


def null_space(A):
    """FieldArray.null_space (_fields/_array.py:1705-1755)."""
    if not A.ndim == 2:
        raise ValueError(f"Only 2-D matrices have a null space, not {A.ndim}-D.")
    return left_null_space(A.T)


# ------------------------------------------------------------------------------------------------
# np.dot / np.vdot / np.inner / np.outer (dot_jit, vdot_jit, inner_jit, outer_jit: _domains/_linalg.py:81-188)
# ------------------------------------------------------------------------------------------------
def _check_pair(field, a, b):
    if not isinstance(a, field):
        raise TypeError(f"Argument 'a' must be a {field.name} array, not {type(a)}.")
    if not isinstance(b, field):
        raise TypeError(f"Argument 'b' must be a {field.name} array, not {type(b)}.")


def dot(field, a, b, out=None):
    _check_pair(field, a, b)
    if out is not None:
        raise NotImplementedError("The 'out' argument is not supported by the CUDA backend's dot().")
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    if a.ndim == 1 and b.ndim == 1:
        return np.sum(a * b)
    if a.ndim == 2 and b.ndim == 2:
        return matmul(field, a, b)
    if a.ndim >= 2 and b.ndim == 1:
        return np.sum(a * b, axis=-1)
    raise NotImplementedError(
        "Currently 'dot' is only supported up to 2-D matrices. "
        "Please open a GitHub issue at https://github.com/mhostetter/galois/issues."
    )


def vdot(field, a, b):
    _check_pair(field, a, b)
    a = a.flatten()
    b = b.flatten().reshape(a.shape)  # mimics NumPy's error scenarios, as the reference does
    return np.sum(a * b)


def inner(field, a, b):
    _check_pair(field, a, b)
    if a.ndim == 0 or b.ndim == 0:
        return a * b
    if not a.shape[-1] == b.shape[-1]:
        raise ValueError(
            f"Operation 'inner' requires 'a' and 'b' to have the same last dimension, not {a.shape} and {b.shape}."
        )
    return np.sum(a * b, axis=-1)


def outer(field, a, b, out=None):
    _check_pair(field, a, b)
    if out is not None:
        raise NotImplementedError("The 'out' argument is not supported by the CUDA backend's outer().")
    return np.multiply.outer(a.ravel(), b.ravel())


# ------------------------------------------------------------------------------------------------
# device-resident entry points (no host round trip of the matrix / roots)
# ------------------------------------------------------------------------------------------------
def row_reduce_device(A, ncols=None, out=None):
    """
    Reduced row echelon form of a device-resident 2-D ``DeviceArray`` (gfb_row_reduce).  ``out`` (same shape/dtype)
    receives the result; by default a copy of ``A`` is reduced.  Returns ``(out, info, det)`` where ``info`` is a CUDA
    int64 tensor ``[n_pivots, n_row_exchanges]`` and ``det`` a one-element DeviceArray with the signed pivot product.
    """
    import torch

    from ._field import DeviceArray

    field = A.field
    _no_wide(field, "row_reduce")
    if len(A.shape) != 2:
        raise ValueError(f"Only 2-D matrices can be converted to reduced row echelon form, not {len(A.shape)}-D.")
    rows, cols = A.shape
    ncols = cols if ncols is None else int(ncols)
    if out is None:
        out = DeviceArray(field, A.tensor.clone(), A.shape, A.dtype)
    elif out is not A:
        out.tensor.copy_(A.tensor)
    lib = _lib.lib()
    ws = torch.empty(lib.gfb_row_reduce_workspace_bytes(field._handle(), rows, cols), dtype=torch.uint8, device="cuda")
    info = torch.zeros(2, dtype=torch.int64, device="cuda")
    det = DeviceArray.empty(field, (1,), A.dtype)
    st = lib.gfb_row_reduce(field._handle(), A.elem_bytes, out.data_ptr, rows, cols, ncols, info.data_ptr(), det.data_ptr,
                            ws.data_ptr(), DeviceArray._stream())
    _lib.check(st, "gfb_row_reduce")
    return out, info, det


def poly_roots_device(nonzero_degrees, nonzero_coeffs, max_roots=None):
    """
    gfb_poly_roots with device-resident coefficients (a 1-D ``DeviceArray``) -> (roots DeviceArray of capacity
    ``max_roots``, CUDA uint64-as-int64 tensor with the number of roots).
    """
    import torch

    from ._field import DeviceArray

    field = nonzero_coeffs.field
    degrees = np.ascontiguousarray(nonzero_degrees, dtype=np.int64).reshape(-1)
    if degrees.size == 0 or degrees.size != nonzero_coeffs.size:
        raise ValueError("Arguments 'nonzero_degrees' and 'nonzero_coeffs' must be non-empty and have the same size.")
    if np.any(np.diff(degrees) >= 0):
        raise ValueError("Argument 'nonzero_degrees' must be strictly descending.")
    max_roots = int(min(int(degrees[0]), field.order)) if max_roots is None else int(max_roots)
    lib = _lib.lib()
    ws = torch.empty(lib.gfb_poly_roots_workspace_bytes(field._handle(), degrees.size), dtype=torch.uint8, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    roots = DeviceArray.empty(field, (max(max_roots, 1),), nonzero_coeffs.dtype)
    st = lib.gfb_poly_roots(field._handle(), nonzero_coeffs.elem_bytes, degrees.ctypes.data, nonzero_coeffs.data_ptr,
                            degrees.size, roots.data_ptr, max_roots, count.data_ptr(), ws.data_ptr(), DeviceArray._stream())
    _lib.check(st, "gfb_poly_roots")
    return roots, count
