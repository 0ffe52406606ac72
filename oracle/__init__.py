"""
oracle — CPU restatement of the reference's hot path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``galois_b200``) never does.

The arithmetic lives in ``gf_oracle.c`` (compiled by ``galois_b200.build.build_oracle`` into
``oracle/_build/libgforacle.so``); this module is the ctypes/NumPy face of it.  Every function works
on int64 arrays like the reference's Numba kernels do.  Parity status: PINNED — checked against the
reference's own Sage/SymPy vectors and against outputs of the live reference (tests/golden/,
tests/test_oracle_golden.py).
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libgforacle.so"

ADD, SUB, MUL, DIV = 0, 1, 2, 3
NEG, INV, SQUARE = 0, 1, 2

_lib = None


class _CField(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int),
        ("m", ctypes.c_int),
        ("p", ctypes.c_int64),
        ("order", ctypes.c_int64),
        ("poly", ctypes.c_int64),
        ("EXP", ctypes.c_void_p),
        ("LOG", ctypes.c_void_p),
    ]


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            from galois_b200.build import build_oracle

            build_oracle()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.orc_max_threads.restype = ctypes.c_int
        _lib.orc_berlekamp_massey.restype = ctypes.c_int64
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _i64(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x), dtype=np.int64)


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


class ZeroDivision(ZeroDivisionError):
    pass


_ZERO_MSG = "Cannot compute the multiplicative inverse of 0 in a Galois field."


class Field:
    """
    A field descriptor for the oracle.

    ``Field.prime(p)``                     GF(p) in the reference's ``jit-calculate`` mode
    ``Field.binary(m, poly)``              GF(2^m) ``jit-calculate`` (shift-and-xor)
    ``Field.binary(m, poly, alpha=...)``   GF(2^m) ``jit-lookup`` with EXP/LOG tables built from ``alpha``
    """

    def __init__(self, kind: int, p: int, m: int, poly: int, alpha: int | None = None):
        self.kind, self.p, self.m, self.poly, self.alpha = kind, p, m, poly, alpha
        self.order = p**m
        self.EXP = self.LOG = None
        if kind == 3:
            self.EXP = np.zeros(2 * self.order, dtype=np.int64)
            self.LOG = np.zeros(self.order, dtype=np.int64)
            rc = lib().orc_build_tables(
                ctypes.c_int(m), ctypes.c_int64(poly), ctypes.c_int64(alpha), _ptr(self.EXP), _ptr(self.LOG)
            )
            if rc != 0:
                raise RuntimeError("primitive element is not a generator")
        self._c = _CField(
            kind,
            m,
            p,
            self.order if self.order < 2**63 else 0,
            poly if poly < 2**63 else 0,
            self.EXP.ctypes.data if self.EXP is not None else None,
            self.LOG.ctypes.data if self.LOG is not None else None,
        )

    @classmethod
    def prime(cls, p: int) -> "Field":
        return cls(1, p, 1, 0)

    @classmethod
    def binary(cls, m: int, poly: int, alpha: int | None = None) -> "Field":
        return cls(3 if alpha is not None else 2, 2, m, poly, alpha)

    @property
    def ref(self):
        return ctypes.byref(self._c)

    # ---------------------------------------------------------------- element-wise
    def binary_op(self, op: int, a, b, threads: int = 1) -> np.ndarray:
        a, b = np.broadcast_arrays(_i64(a), _i64(b))
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        out = np.empty_like(a)
        rc = lib().orc_binary(self.ref, op, _ptr(a), _ptr(b), _ptr(out), ctypes.c_int64(a.size), threads)
        if rc:
            raise ZeroDivision(_ZERO_MSG)
        return out

    def unary_op(self, op: int, a, threads: int = 1) -> np.ndarray:
        a = _i64(a)
        out = np.empty_like(a)
        rc = lib().orc_unary(self.ref, op, _ptr(a), _ptr(out), ctypes.c_int64(a.size), threads)
        if rc:
            raise ZeroDivision(_ZERO_MSG)
        return out

    def add(self, a, b, **kw):
        return self.binary_op(ADD, a, b, **kw)

    def sub(self, a, b, **kw):
        return self.binary_op(SUB, a, b, **kw)

    def mul(self, a, b, **kw):
        return self.binary_op(MUL, a, b, **kw)

    def div(self, a, b, **kw):
        return self.binary_op(DIV, a, b, **kw)

    def neg(self, a, **kw):
        return self.unary_op(NEG, a, **kw)

    def inv(self, a, **kw):
        return self.unary_op(INV, a, **kw)

    def square(self, a, **kw):
        return self.unary_op(SQUARE, a, **kw)

    def power(self, a, e, threads: int = 1) -> np.ndarray:
        a, e = np.broadcast_arrays(_i64(a), _i64(e))
        a, e = np.ascontiguousarray(a), np.ascontiguousarray(e)
        out = np.empty_like(a)
        rc = lib().orc_power(self.ref, _ptr(a), _ptr(e), _ptr(out), ctypes.c_int64(a.size), threads)
        if rc:
            raise ZeroDivision(_ZERO_MSG)
        return out

    # ---------------------------------------------------------------- transforms
    def fft(self, x, omega: int, scale: int = 1, threads: int = 1) -> np.ndarray:
        """1-D or batched ([batch, N]) DFT with the reference's iterative mixed-radix algorithm."""
        x = _i64(x)
        squeeze = x.ndim == 1
        x2 = np.atleast_2d(x)
        batch, N = x2.shape
        factors = np.array(prime_factors(N), dtype=np.int64)
        out = np.empty_like(x2)
        lib().orc_fft_batch(
            self.ref,
            _ptr(x2),
            ctypes.c_int64(batch),
            ctypes.c_int64(N),
            ctypes.c_int64(omega),
            _ptr(factors),
            ctypes.c_int(len(factors)),
            ctypes.c_int64(scale),
            _ptr(out),
            threads,
        )
        return out[0] if squeeze else out

    def convolve(self, a, b) -> np.ndarray:
        a, b = _i64(a), _i64(b)
        out = np.empty(a.size + b.size - 1, dtype=np.int64)
        lib().orc_convolve(self.ref, _ptr(a), ctypes.c_int64(a.size), _ptr(b), ctypes.c_int64(b.size), _ptr(out))
        return out

    def matmul(self, A, B, threads: int = 1) -> np.ndarray:
        A, B = _i64(A), _i64(B)
        A3 = A.reshape((-1,) + A.shape[-2:])
        b_batched = B.ndim > 2
        B3 = B.reshape((-1,) + B.shape[-2:])
        batch, M, K = A3.shape
        N = B3.shape[-1]
        C = np.empty((batch, M, N), dtype=np.int64)
        lib().orc_matmul(
            self.ref,
            _ptr(A3),
            _ptr(B3),
            _ptr(C),
            ctypes.c_int64(batch),
            ctypes.c_int64(M),
            ctypes.c_int64(K),
            ctypes.c_int64(N),
            ctypes.c_int(1 if b_batched else 0),
            threads,
        )
        return C.reshape(A.shape[:-2] + (M, N))

    def poly_evaluate(self, coeffs_desc, values) -> np.ndarray:
        c, v = _i64(coeffs_desc), _i64(values)
        y = np.empty_like(v)
        lib().orc_poly_evaluate(self.ref, _ptr(c), ctypes.c_int64(c.size), _ptr(v), ctypes.c_int64(v.size), _ptr(y))
        return y

    def berlekamp_massey(self, seq) -> np.ndarray:
        s = _i64(seq)
        out = np.zeros(max(s.size, 1), dtype=np.int64)
        n = lib().orc_berlekamp_massey(self.ref, _ptr(s), ctypes.c_int64(s.size), _ptr(out))
        return out[:n]


    # ---------------------------------------------------------------- polynomial roots
    def poly_roots(self, nonzero_degrees, nonzero_coeffs, primitive_element: int | None = None) -> np.ndarray:
        """roots_jit.__call__ (_polys/_dense.py:448-466): Chien search, then the roots in increasing order."""
        d, c = _i64(nonzero_degrees), _i64(nonzero_coeffs)
        alpha = self.alpha if primitive_element is None else primitive_element
        cap = int(min(max(int(d.max()), 1) + 2, self.order))
        if not c.any():
            cap = self.order  # the zero polynomial: every power of alpha passes the test
        roots = np.zeros(cap, dtype=np.int64)
        powers = np.zeros(cap, dtype=np.int64)
        lib().orc_poly_roots.restype = ctypes.c_int64
        n = lib().orc_poly_roots(
            self.ref, _ptr(d), _ptr(c), ctypes.c_int64(d.size), ctypes.c_int64(alpha), _ptr(roots), _ptr(powers)
        )
        return np.sort(roots[:n])

    # ---------------------------------------------------------------- dense linear algebra
    def row_reduce(self, A, ncols: int | None = None):
        """row_reduce_jit.__call__ (_domains/_linalg.py:321-353) -> (A_rre, n_pivots)."""
        A = _i64(A).copy()
        rows, cols = A.shape
        ncols = cols if ncols is None else ncols
        lib().orc_row_reduce.restype = ctypes.c_int64
        p = lib().orc_row_reduce(self.ref, _ptr(A), ctypes.c_int64(rows), ctypes.c_int64(cols), ctypes.c_int64(ncols))
        return A, int(p)

    def plu_decompose(self, A):
        """plu_decompose_jit.__call__ (_linalg.py:390-428) -> (P, L, U, n_permutations) with A = P L U."""
        U = _i64(A).copy()
        m, n = U.shape
        L = np.zeros((m, m), dtype=np.int64)
        perm = np.zeros(m, dtype=np.int64)
        lib().orc_plu_decompose.restype = ctypes.c_int64
        n_perm = lib().orc_plu_decompose(self.ref, _ptr(U), _ptr(L), _ptr(perm), ctypes.c_int64(m), ctypes.c_int64(n))
        P_rows = np.zeros((m, m), dtype=np.int64)  # the reference's row-swapped identity
        P_rows[np.arange(m), perm] = 1
        return P_rows.T.copy(), L, U, int(n_perm)

    def lu_decompose(self, A):
        """lu_decompose_jit.__call__ (_linalg.py:356-387) -> (L, U); ValueError when a row swap would be needed."""
        U = _i64(A).copy()
        m, n = U.shape
        L = np.zeros((m, m), dtype=np.int64)
        rc = lib().orc_lu_decompose(self.ref, _ptr(U), _ptr(L), ctypes.c_int64(m), ctypes.c_int64(n))
        if rc:
            raise ValueError("The LU decomposition of 'A' does not exist. Use the PLU decomposition instead.")
        return L, U

    def matrix_rank(self, A) -> int:
        """matrix_rank_jit (_linalg.py:479-489)."""
        R, _ = self.row_reduce(A)
        return int(np.sum(~np.all(R == 0, axis=1)))

    def row_reduce_eye(self, A, ncols: int | None = None, eye: str = "left"):
        """FieldArray.row_reduce (_fields/_array.py:1412-1469): eye="right" reduces the doubly-reversed matrix."""
        A = _i64(A)
        if eye == "left":
            return self.row_reduce(A, ncols)[0]
        return self.row_reduce(A[::-1, ::-1], ncols)[0][::-1, ::-1].copy()

    def row_space(self, A):
        """FieldArray.row_space (_fields/_array.py:1541-1590)."""
        R, _ = self.row_reduce(A)
        rank = int(np.sum(~np.all(R == 0, axis=1)))
        return R[:rank, :]

    def column_space(self, A):
        return self.row_space(_i64(A).T)

    def left_null_space(self, A):
        """FieldArray.left_null_space (_fields/_array.py:1639-1703)."""
        A = _i64(A)
        m, n = A.shape
        AI = np.concatenate([A, np.eye(m, dtype=np.int64)], axis=-1)
        R, p = self.row_reduce(AI, ncols=n)
        LN = R[p:, n:]
        if LN.shape[0] == 0:
            return LN.copy()
        return self.row_reduce(LN)[0]

    def null_space(self, A):
        return self.left_null_space(_i64(A).T)

    def matrix_inv(self, A):
        """inv_jit (_linalg.py:492-522): row-reduce [A | I] on the first n columns."""
        A = _i64(A)
        n = A.shape[0]
        AI = np.concatenate([A, np.eye(n, dtype=np.int64)], axis=-1)
        R, _ = self.row_reduce(AI, ncols=n)
        rank = int(np.sum(~np.all(R[:, :n] == 0, axis=1)))
        if rank != n:
            raise np.linalg.LinAlgError(
                f"Argument 'A' is singular and not invertible because it does not have full rank of {n}, "
                f"but rank of {rank}."
            )
        return R[:, -n:].copy()

    def solve(self, A, b):
        """solve_jit (_linalg.py:525-546): x = A^-1 b."""
        b = _i64(b)
        Ainv = self.matrix_inv(A)
        if b.ndim == 1:
            return self.matmul(Ainv, b.reshape(-1, 1)).reshape(-1)
        return self.matmul(Ainv, b)

    def det(self, A) -> int:
        """det_jit (_linalg.py:446-471): closed forms for n = 2, 3, otherwise (-1)^swaps det(L) det(U)."""
        A = _i64(A)
        n = A.shape[0]
        mul, sub = (lambda x, y: int(self.mul(x, y).item())), (lambda x, y: int(self.sub(x, y).item()))
        if n == 2:
            return sub(mul(A[0, 0], A[1, 1]), mul(A[0, 1], A[1, 0]))
        if n == 3:
            t0 = mul(A[0, 0], sub(mul(A[1, 1], A[2, 2]), mul(A[1, 2], A[2, 1])))
            t1 = mul(A[0, 1], sub(mul(A[1, 0], A[2, 2]), mul(A[1, 2], A[2, 0])))
            t2 = mul(A[0, 2], sub(mul(A[1, 0], A[2, 1]), mul(A[1, 1], A[2, 0])))
            return int(self.add(sub(t0, t1), t2).item())
        _, L, U, n_perm = self.plu_decompose(A)
        det = 1
        for i in range(n):
            det = mul(det, mul(L[i, i], U[i, i]))
        return int(self.neg(det).item()) if n_perm % 2 else det


def prime_factors(n: int) -> list[int]:
    """Prime factors with multiplicity, ascending (fft_jit._prime_factors, _function.py:214-229)."""
    out, d = [], 2
    while n > 1 and d * d <= n:
        while n % d == 0:
            out.append(d)
            n //= d
        d += 1
    if n > 1:
        out.append(n)
    return out


class ReedSolomon:
    """RS(n, k) over ``field`` with roots alpha^c..alpha^(c+d-2) (restates _reed_solomon.py:111-218)."""

    def __init__(self, field: Field, n: int, k: int, c: int = 1, alpha: int = 2):
        self.field, self.n, self.k, self.c, self.alpha = field, n, k, c, alpha
        self.d = n - k + 1
        self.roots = np.zeros(max(self.d - 1, 1), dtype=np.int64)
        self.generator_poly = np.zeros(self.d, dtype=np.int64)
        lib().orc_rs_generator_poly(
            field.ref, n, k, c, ctypes.c_int64(alpha), _ptr(self.roots), _ptr(self.generator_poly)
        )
        self.roots = self.roots[: self.d - 1]
        self.P = np.zeros((k, n - k), dtype=np.int64)
        lib().orc_rs_parity_matrix(field.ref, n, k, _ptr(self.generator_poly), _ptr(self.P))

    def encode(self, msg, threads: int = 1) -> np.ndarray:
        msg = np.atleast_2d(_i64(msg))
        N, ks = msg.shape
        cw = np.empty((N, ks + self.n - self.k), dtype=np.int64)
        lib().orc_rs_encode(
            self.field.ref, self.n, self.k, _ptr(self.P), _ptr(msg), ks, _ptr(cw), ctypes.c_int64(N), threads
        )
        return cw

    def decode(self, codewords, erasures=None, threads: int = 1):
        """Returns (decoded codewords [N, ns] int64, n_errors [N] int64)."""
        cw = np.atleast_2d(_i64(codewords))
        N, ns = cw.shape
        er = None
        if erasures is not None:
            er = np.ascontiguousarray(np.atleast_2d(np.asarray(erasures)).astype(np.uint8))
        dec = np.empty_like(cw)
        n_err = np.zeros(N, dtype=np.int64)
        rc = lib().orc_rs_decode(
            self.field.ref,
            _ptr(cw),
            _ptr(er) if er is not None else None,
            ctypes.c_int64(N),
            ns,
            self.n,
            ctypes.c_int64(self.alpha),
            self.c,
            _ptr(self.roots),
            self.d,
            _ptr(dec),
            _ptr(n_err),
            threads,
        )
        if rc:
            raise ValueError("oracle rs_decode: bad arguments")
        return dec, n_err


# ------------------------------------------------------------------------ GF(2^m), 64 < m <= 128
def _limbs(x) -> np.ndarray:
    """[n, 2] uint64 (lo, hi) from an array of Python ints or an existing limb array."""
    if isinstance(x, np.ndarray) and x.dtype == np.uint64 and x.ndim == 2 and x.shape[1] == 2:
        return np.ascontiguousarray(x)
    vals = [int(v) for v in np.asarray(x, dtype=object).ravel()]
    out = np.empty((len(vals), 2), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        out[i, 0] = v & mask
        out[i, 1] = v >> 64
    return out


def limbs_to_ints(limbs: np.ndarray) -> list[int]:
    return [int(lo) | (int(hi) << 64) for lo, hi in limbs]


class Field128:
    """GF(2^m), 64 < m <= 128.  ``poly`` is the full irreducible polynomial as a Python int."""

    def __init__(self, m: int, poly: int):
        self.m, self.poly = m, poly
        low = poly & ((1 << 128) - 1)
        self.poly_lo, self.poly_hi = low & ((1 << 64) - 1), low >> 64

    def binary_op(self, op: int, a, b, threads: int = 1) -> np.ndarray:
        a, b = _limbs(a), _limbs(b)
        out = np.empty_like(a)
        rc = lib().orc128_binary(
            self.m,
            ctypes.c_uint64(self.poly_lo),
            ctypes.c_uint64(self.poly_hi),
            op,
            _ptr(a),
            _ptr(b),
            _ptr(out),
            ctypes.c_int64(a.shape[0]),
            threads,
        )
        if rc:
            raise ZeroDivision(_ZERO_MSG)
        return out

    def unary_op(self, op: int, a, threads: int = 1) -> np.ndarray:
        a = _limbs(a)
        out = np.empty_like(a)
        rc = lib().orc128_unary(
            self.m,
            ctypes.c_uint64(self.poly_lo),
            ctypes.c_uint64(self.poly_hi),
            op,
            _ptr(a),
            _ptr(out),
            ctypes.c_int64(a.shape[0]),
            threads,
        )
        if rc:
            raise ZeroDivision(_ZERO_MSG)
        return out

    def power(self, a, e, threads: int = 1) -> np.ndarray:
        a = _limbs(a)
        e = np.ascontiguousarray(np.broadcast_to(_i64(e), (a.shape[0],)))
        out = np.empty_like(a)
        rc = lib().orc128_power(
            self.m,
            ctypes.c_uint64(self.poly_lo),
            ctypes.c_uint64(self.poly_hi),
            _ptr(a),
            _ptr(e),
            _ptr(out),
            ctypes.c_int64(a.shape[0]),
            threads,
        )
        if rc:
            raise ZeroDivision(_ZERO_MSG)
        return out

    def mul(self, a, b, **kw):
        return self.binary_op(MUL, a, b, **kw)

    def div(self, a, b, **kw):
        return self.binary_op(DIV, a, b, **kw)

    def add(self, a, b, **kw):
        return self.binary_op(ADD, a, b, **kw)

    def inv(self, a, **kw):
        return self.unary_op(INV, a, **kw)


def prime_mul_i64(p: int, a: np.ndarray, b: np.ndarray, out: np.ndarray, threads: int = 1) -> None:
    """The reference's int64 GF(p) multiply loop, for the CPU baseline (no allocation)."""
    lib().orc_prime_mul_i64(ctypes.c_int64(p), _ptr(a), _ptr(b), _ptr(out), ctypes.c_int64(a.size), threads)


# ------------------------------------------------------------------------ binary BCH
def _gf2m_mul(a: int, b: int, m: int, poly: int) -> int:
    """multiply_binary.calculate, _domains/_calculate.py:309-324 (pure Python; cold path)."""
    c, order = 0, 1 << m
    while b:
        if b & 1:
            c ^= a
        b >>= 1
        a <<= 1
        if a >= order:
            a ^= poly
    return c


def _gf2m_pow(a: int, e: int, m: int, poly: int) -> int:
    r = 1
    while e:
        if e & 1:
            r = _gf2m_mul(r, a, m, poly)
        a = _gf2m_mul(a, a, m, poly)
        e >>= 1
    return r


def _gf2_polymul(a: int, b: int) -> int:
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        b >>= 1
    return r


def bch_generator_poly(m: int, poly: int, d: int, alpha: int, c: int):
    """_generator_poly_from_d, _codes/_bch.py:1178-1197, for field = GF(2): g(x) as a GF(2)[x] bit mask and the
    d-1 roots alpha^(c+i).  The minimal polynomial of a root is prod (x - root^(2^j)) over its conjugates."""
    g, roots, seen = 1, [], []
    for i in range(d - 1):
        root = _gf2m_pow(alpha, c + i, m, poly)
        roots.append(root)
        conj, b = [], root
        while b not in conj:
            conj.append(b)
            b = _gf2m_mul(b, b, m, poly)
        coeffs = [1]
        for cj in conj:
            nxt = [0] * (len(coeffs) + 1)
            for j, a in enumerate(coeffs):
                nxt[j + 1] ^= a
                nxt[j] ^= _gf2m_mul(a, cj, m, poly)
            coeffs = nxt
        assert all(v in (0, 1) for v in coeffs)
        mi = sum(v << j for j, v in enumerate(coeffs))
        if mi not in seen:
            g = _gf2_polymul(g, mi)
            seen.append(mi)
    return g, roots


class BCH:
    """Binary BCH(n, k, d): symbols in GF(2), decoding arithmetic in ``ext`` = GF(2^m) (a lookup ``Field``).
    Encode is the systematic product over GF(2) (_linear.py:270-284), decode is bch_decode_jit with
    field = GF(2), extension_field = GF(2^m) (_bch.py:721-727, 1337-1578)."""

    def __init__(self, ext: Field, n: int, d: int, c: int = 1, alpha: int = 2):
        self.ext, self.n, self.d, self.c, self.alpha = ext, n, d, c, alpha
        g, roots = bch_generator_poly(ext.m, ext.poly, d, alpha, c)
        r = g.bit_length() - 1
        self.k = n - r
        self.generator_poly = np.array([(g >> (r - i)) & 1 for i in range(r + 1)], dtype=np.int64)
        self.roots = np.array(roots, dtype=np.int64)
        self.base = Field.prime(2)
        self.P = np.zeros((self.k, max(r, 1)), dtype=np.int64)
        if r:
            lib().orc_rs_parity_matrix(self.base.ref, n, self.k, _ptr(self.generator_poly), _ptr(self.P))
        self.P = self.P[:, :r]

    def encode(self, msg, threads: int = 1) -> np.ndarray:
        msg = np.atleast_2d(_i64(msg))
        N, ks = msg.shape
        cw = np.empty((N, ks + self.n - self.k), dtype=np.int64)
        P = np.ascontiguousarray(self.P)
        lib().orc_rs_encode(self.base.ref, self.n, self.k, _ptr(P), _ptr(msg), ks, _ptr(cw), ctypes.c_int64(N), threads)
        return cw

    def decode(self, codewords, erasures=None, threads: int = 1):
        cw = np.atleast_2d(_i64(codewords))
        N, ns = cw.shape
        er = None
        if erasures is not None:
            er = np.ascontiguousarray(np.atleast_2d(np.asarray(erasures)).astype(np.uint8))
        dec = np.empty_like(cw)
        n_err = np.zeros(N, dtype=np.int64)
        rc = lib().orc_bch_decode(
            self.ext.ref, self.base.ref, _ptr(cw), _ptr(er) if er is not None else None, ctypes.c_int64(N), ns, self.n,
            ctypes.c_int64(self.alpha), self.c, _ptr(self.roots), self.d, _ptr(dec), _ptr(n_err), threads)
        if rc:
            raise ValueError("oracle bch_decode: bad arguments")
        # dec_codeword.astype(codeword.dtype), _bch.py:1299: GF(2) arrays are uint8
        return dec.astype(np.uint8).astype(np.int64), n_err
