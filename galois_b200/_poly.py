"""
Polynomial evaluation and root finding — the two array kernels behind ``galois.Poly.__call__`` and
``galois.Poly.roots``, plus a slim ``Poly`` holder that gives them the reference's call shape.

    poly_evaluate(coeffs, x)                    evaluate_elementwise_jit.__call__  (_polys/_dense.py:409-424)
    poly_roots(nonzero_degrees, nonzero_coeffs) roots_jit.__call__                 (_polys/_dense.py:448-466)
    Poly(coeffs, field=GF)(x), .roots()         Poly.__call__ / Poly.roots         (_polys/_poly.py:711-790, 1028-1100)

Polynomial ALGEBRA (add / mul / divmod / gcd / factoring, and the root multiplicities that need it) is the reference's
host-side control code around these kernels and is not rebuilt here.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, _pinned
from ._field import FieldArray
from ._linalg import poly_evaluate  # noqa: F401  (re-exported)


def poly_roots(nonzero_degrees, nonzero_coeffs):
    """
    Roots of ``sum_j nonzero_coeffs[j] * x**nonzero_degrees[j]`` in the field of ``nonzero_coeffs``, in increasing
    order, with the coefficients' dtype.  Mirrors ``roots_jit.__call__`` (_polys/_dense.py:448-466).
    """
    if not isinstance(nonzero_degrees, np.ndarray):
        raise TypeError(f"Argument 'nonzero_degrees' must be a np.ndarray, not {type(nonzero_degrees)}.")
    if not isinstance(nonzero_coeffs, FieldArray):
        raise TypeError(f"Argument 'nonzero_coeffs' must be a FieldArray, not {type(nonzero_coeffs)}.")
    field = type(nonzero_coeffs)
    degrees = np.ascontiguousarray(nonzero_degrees, dtype=np.int64).reshape(-1)
    if degrees.size == 0:
        # the zero polynomial: np.max of an empty array inside the reference kernel (_dense.py:485)
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    if degrees.size != nonzero_coeffs.size:
        raise ValueError("Arguments 'nonzero_degrees' and 'nonzero_coeffs' must have the same size.")
    if field._wide or field.order > 2**33:
        raise NotImplementedError(
            f"Root finding is an exhaustive search over the field; {field.name} has too many elements for it."
        )
    dtype = nonzero_coeffs.dtype
    order_idx = np.argsort(-degrees, kind="stable")
    degrees = np.ascontiguousarray(degrees[order_idx])
    cv = field._to_dev_layout(field._prep(nonzero_coeffs.reshape(-1)[order_idx].view(field), dtype))
    max_roots = int(min(int(degrees[0]), field.order))
    out = _pinned.empty(max(max_roots, 1), dtype)
    n_roots = ctypes.c_size_t(0)
    st = _lib.lib().gfb_poly_roots_host(field._handle(), field._elem_bytes(dtype), degrees.ctypes.data, cv.ctypes.data,
                                        degrees.size, out.ctypes.data, max_roots, ctypes.byref(n_roots))
    _lib.check(st, "gfb_poly_roots_host")
    n = min(int(n_roots.value), max_roots)
    return np.array(out[:n], dtype=dtype).view(field)


def _coeffs_from_ints(field, coeffs):
    """Integer coefficients -> field array; like galois.Poly, a negative integer -c denotes -field(c) (_poly.py:126-131)."""
    arr = np.atleast_1d(np.asarray(coeffs, dtype=object if field._wide else np.int64))
    neg = arr < 0
    if neg.any():
        p = field.characteristic
        mag = np.abs(arr)
        field(mag)  # range check of the magnitudes
        # additive inverse of a literal: p - c in GF(p), c itself in characteristic 2 (construction-time constant
        # folding, not array arithmetic)
        arr = np.where(neg, mag if p == 2 else (p - mag) % p, arr)
    return field(arr)


class Poly:
    """
    A univariate polynomial over a field of this package, stored sparsely (non-zero degrees, descending, with their
    coefficients) — the slice of ``galois.Poly`` the evaluation and root kernels are called through.
    """

    __slots__ = ("_field", "_degrees", "_coeffs")

    def __init__(self, coeffs, field=None, order: str = "desc"):
        if order not in ("desc", "asc"):
            raise ValueError(f"Argument 'order' must be in ['desc', 'asc'], not {order!r}.")
        if isinstance(coeffs, FieldArray):
            if field is not None and field is not type(coeffs):
                raise TypeError("Argument 'field' must match the field of 'coeffs'.")
            field = type(coeffs)
            c = coeffs
        else:
            if field is None:
                from ._field import GF

                field = GF(2)
            # like galois.Poly, negative integers denote additive inverses: -c -> -field(c)  (_poly.py:126-131)
            c = _coeffs_from_ints(field, coeffs)
        c = np.atleast_1d(c)
        if c.ndim != 1:
            raise ValueError(f"Argument 'coeffs' can have dimension at most 1, not {c.ndim}.")
        if order == "asc":
            c = c[::-1]
        raw = np.asarray(c)
        nz = np.nonzero(raw)[0]
        self._field = field
        self._degrees = (raw.size - 1 - nz).astype(np.int64)
        self._coeffs = c[nz].view(field)

    @classmethod
    def Degrees(cls, degrees, coeffs=None, field=None):  # noqa: N802
        """Poly.Degrees (_polys/_poly.py:426-520): a polynomial from its non-zero degrees and coefficients."""
        degrees = np.asarray(degrees, dtype=np.int64).reshape(-1)
        if degrees.size and degrees.min() < 0:
            raise ValueError(f"Argument 'degrees' must have non-negative values, not {degrees}.")
        if np.unique(degrees).size != degrees.size:
            raise ValueError(f"Argument 'degrees' must have unique entries, not {degrees}.")
        if coeffs is None:
            coeffs = [1] * degrees.size
        if isinstance(coeffs, FieldArray):
            field = type(coeffs)
            c = np.atleast_1d(coeffs)
        else:
            if field is None:
                from ._field import GF

                field = GF(2)
            c = _coeffs_from_ints(field, coeffs)
        if c.size != degrees.size:
            raise ValueError("Arguments 'degrees' and 'coeffs' must have the same length.")
        idx = np.argsort(-degrees, kind="stable")
        keep = idx[np.asarray(c)[idx] != 0]
        self = cls.__new__(cls)
        self._field = field
        self._degrees = degrees[keep]
        self._coeffs = c[keep].view(field)
        return self

    # ------------------------------------------------------------------ properties (names as in galois.Poly)
    @property
    def field(self):
        return self._field

    @property
    def degree(self) -> int:
        return int(self._degrees[0]) if self._degrees.size else 0

    @property
    def nonzero_degrees(self) -> np.ndarray:
        return self._degrees.copy()

    @property
    def nonzero_coeffs(self):
        return self._coeffs.copy()

    @property
    def coeffs(self):
        """Dense coefficients, degree-descending."""
        if self.degree + 1 > 1 << 26:
            raise MemoryError("dense coefficients of a very high-degree sparse polynomial")
        dtype = self._coeffs.dtype if self._coeffs.size else self._field.dtypes[0]
        dense = self._field.Zeros(self.degree + 1, dtype=dtype)
        if self._degrees.size:
            np.asarray(dense)[self.degree - self._degrees] = np.asarray(self._coeffs)
        return dense

    def __eq__(self, other):
        if isinstance(other, Poly):
            return (self._field is other._field and np.array_equal(self._degrees, other._degrees)
                    and np.array_equal(np.asarray(self._coeffs), np.asarray(other._coeffs)))
        if isinstance(other, (int, np.integer)):
            return np.array_equal(np.asarray(self.coeffs), [int(other)])
        return NotImplemented

    def __repr__(self):
        terms = []
        for d, c in zip(self._degrees.tolist(), np.asarray(self._coeffs).tolist()):
            mono = "" if d == 0 else ("x" if d == 1 else f"x^{d}")
            terms.append(mono if (c == 1 and mono) else f"{c}{mono}")
        return f"Poly({' + '.join(terms) if terms else '0'}, {self._field.name})"

    # ------------------------------------------------------------------ the two kernels
    def __call__(self, at):
        """Evaluate element-wise (Poly.__call__ with elementwise=True, _polys/_poly.py:1028-1100)."""
        field = self._field
        if isinstance(at, (int, np.integer)):
            x = field(int(at))
        elif isinstance(at, field):
            x = at
        elif isinstance(at, FieldArray):
            raise TypeError(f"Argument 'at' must be an array over {field.name}, not {type(at).name}.")
        else:
            x = field(at)
        return poly_evaluate(self.coeffs, x)

    def roots(self, multiplicity: bool = False):
        """The roots in increasing order (Poly.roots, _polys/_poly.py:711-787)."""
        if not isinstance(multiplicity, bool):
            raise TypeError(f"Argument 'multiplicity' must be a bool, not {type(multiplicity)}.")
        if multiplicity:
            raise NotImplementedError(
                "Root multiplicities need polynomial division (_polys/_poly.py:1672-1698), which is host-side "
                "polynomial algebra outside the CUDA path; use roots() for the distinct roots."
            )
        return poly_roots(self._degrees, self._coeffs)
