"""
``ReedSolomon`` — host-side mirror of galois.ReedSolomon for codes over GF(2^m), m <= 8, whose
encode / decode inner loops run as CUDA kernels (gfb_rs_encode / gfb_rs_decode).

Mirrors /root/reference/src/galois/_codes/_reed_solomon.py:111-218 (constructor arguments and
conventions: alpha = primitive n-th root of unity, roots alpha^c..alpha^(c+d-2), systematic [I | P]),
_codes/_linear.py:58-186 (encode / detect / decode argument handling, shortened codes, 1-D vs 2-D)
and _codes/_cyclic.py:129-138 (message = first ks symbols of a systematic codeword).
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib, _nt, _pinned
from ._field import GF, DeviceArray, FieldArray

# galois.matlab_primitive_poly(2, m) — the default field polynomial of ReedSolomon (_reed_solomon.py:180-185)
_MATLAB_PRIMITIVE_POLY_2 = {1: 0x3, 2: 0x7, 3: 0xB, 4: 0x13, 5: 0x25, 6: 0x43, 7: 0x89, 8: 0x11D, 9: 0x211,
                            10: 0x409, 11: 0x805, 12: 0x1053, 13: 0x201B, 14: 0x4443, 15: 0x8003, 16: 0x1100B}


class ReedSolomon:
    """A general RS(n, k) code over GF(2^m), m <= 8 (systematic)."""

    def __init__(self, n: int, k: int | None = None, d: int | None = None, field=None, alpha=None, c: int = 1,
                 systematic: bool = True):
        for name, v, opt in (("n", n, False), ("k", k, True), ("d", d, True), ("c", c, False)):
            if not (isinstance(v, (int, np.integer)) or (opt and v is None)):
                raise TypeError(f"Argument {name!r} must be an int, not {type(v)}.")
        if not isinstance(systematic, bool):
            raise TypeError(f"Argument 'systematic' must be a bool, not {type(systematic)}.")
        if field is not None and not (isinstance(field, type) and issubclass(field, FieldArray)):
            raise TypeError(f"Argument 'field' must be a FieldArray subclass, not {field}.")
        if d is not None and not d >= 1:
            raise ValueError(f"Argument 'd' must be at least 1, not {d}.")
        if not c >= 0:
            raise ValueError(f"Argument 'c' must be at least 0, not {c}.")
        if not systematic:
            raise NotImplementedError("Non-systematic Reed-Solomon codes are outside the CUDA hot path (SURVEY.md §2.2).")
        if field is None:
            m = _nt.ilog(n, 2) + 1
            field = GF(2**m, irreducible_poly=_MATLAB_PRIMITIVE_POLY_2[m])
        if field.characteristic != 2 or field.degree > 8:
            raise NotImplementedError("The CUDA Reed-Solomon kernels cover codes over GF(2^m), m <= 8.")
        if alpha is None:
            alpha = field.primitive_root_of_unity(n)
        else:
            alpha = field(alpha)
        if d is not None and k is not None:
            if not d == n - k + 1:
                raise ValueError("Arguments 'k' and 'd' were provided but are inconsistent. For Reed-Solomon codes, d = n - k + 1.")
        elif d is not None:
            k = n - (d - 1)
        elif k is not None:
            d = (n - k) + 1
        else:
            raise ValueError("Argument 'k' or 'd' must be provided to define the code size.")
        if not 1 <= k <= n:
            raise ValueError(f"Argument 'k' must be in [1, n], not {k}.")
        if (field.order - 1) % n != 0:
            raise ValueError(f"Argument 'n' must divide q - 1 = {field.order - 1}, not {n}.")

        self._field, self._n, self._k, self._d = field, int(n), int(k), int(d)
        self._alpha, self._c, self._systematic = alpha, int(c), True
        self._is_primitive = n == field.order - 1
        self._is_narrow_sense = c == 1
        self._codes: dict = {}  # device -> gfb_rs_code*
        self._generator_poly = None
        self._P = None

    # ------------------------------------------------------------------------------------ handles
    def _code(self):
        lib = _lib.lib()
        h = self._field._handle()
        code = self._codes.get(h)
        if code is None:
            out = ctypes.c_void_p()
            st = lib.gfb_rs_create(ctypes.byref(out), h, self._n, self._k, self._c, ctypes.c_uint64(int(self._alpha)))
            _lib.check(st, f"gfb_rs_create({self._n},{self._k})")
            code = out.value
            self._codes[h] = code
        return code

    # ------------------------------------------------------------------------------------ properties
    field = property(lambda self: self._field)
    n = property(lambda self: self._n)
    k = property(lambda self: self._k)
    d = property(lambda self: self._d)
    t = property(lambda self: (self._d - 1) // 2)
    alpha = property(lambda self: self._alpha)
    c = property(lambda self: self._c)
    is_systematic = property(lambda self: True)
    is_primitive = property(lambda self: self._is_primitive)
    is_narrow_sense = property(lambda self: self._is_narrow_sense)

    @property
    def roots(self) -> FieldArray:
        a = int(self._alpha)
        poly = int(self._field.irreducible_poly)
        return self._field([_nt.powmod2(a, self._c + i, poly) for i in range(self._d - 1)])

    @property
    def generator_poly(self) -> FieldArray:
        """Coefficients of g(x), degree-descending (the reference returns a Poly; its ``.coeffs`` equal this)."""
        if self._generator_poly is None:
            buf = np.zeros(self._d, dtype=np.uint8)
            _lib.check(_lib.lib().gfb_rs_get_generator_poly(self._code(), buf.ctypes.data), "gfb_rs_get_generator_poly")
            self._generator_poly = self._field(buf)
        return self._generator_poly

    @property
    def G(self) -> FieldArray:
        """Systematic generator matrix [I | P] (_codes/_cyclic.py:206-217)."""
        if self._P is None:
            buf = np.zeros((self._k, self._n - self._k), dtype=np.uint8)
            _lib.check(_lib.lib().gfb_rs_get_parity_matrix(self._code(), buf.ctypes.data), "gfb_rs_get_parity_matrix")
            self._P = buf
        return self._field(np.hstack([np.identity(self._k, dtype=np.uint8), self._P]))

    @property
    def H(self) -> FieldArray:
        """Parity-check matrix: H[i, j] = roots[i]^(n-1-j) (_reed_solomon.py:218)."""
        poly = int(self._field.irreducible_poly)
        rows = [[_nt.powmod2(int(r), self._n - 1 - j, poly) for j in range(self._n)] for r in self.roots]
        return self._field(rows)

    def __repr__(self):
        return f"<Reed-Solomon Code: [{self._n}, {self._k}, {self._d}] over {self._field.name}>"

    # ------------------------------------------------------------------------------------ encode
    def _check_message(self, message):
        message = message if isinstance(message, self._field) else self._field(message)
        if message.ndim > 2:
            raise ValueError(f"Argument 'message' can be either 1-D or 2-D, not {message.ndim}-D.")
        if not 1 <= message.shape[-1] <= self._k:
            raise ValueError(f"Argument 'message' must be a 1-D or 2-D array with last dimension between 1 and {self._k}, not shape {message.shape}.")
        return np.atleast_2d(message), message.ndim == 1

    def encode(self, message, output: str = "codeword") -> FieldArray:
        """_LinearCode.encode (_codes/_linear.py:58-93)."""
        if output not in ("codeword", "parity"):
            raise ValueError(f"Argument 'output' must be in ['codeword', 'parity'], not {output!r}.")
        message, is_1d = self._check_message(message)
        N, ks = message.shape
        r = self._n - self._k
        src = np.ascontiguousarray(message.view(np.ndarray), dtype=np.uint8)
        cw = _pinned.empty((N, ks + r), np.uint8)
        if N:
            st = _lib.lib().gfb_rs_encode_host(self._code(), src.ctypes.data, ks, cw.ctypes.data, N)
            _lib.check(st, "gfb_rs_encode_host")
        codeword = cw.astype(message.dtype, copy=False).view(self._field)
        if output == "parity":
            codeword = codeword[:, ks:]
        return codeword[0, :] if is_1d else codeword

    # ------------------------------------------------------------------------------------ decode
    def _check_codeword(self, codeword):
        codeword = codeword if isinstance(codeword, self._field) else self._field(codeword)
        lo = self._n - self._k + 1
        if not lo <= codeword.shape[-1] <= self._n:
            raise ValueError(f"Argument 'codeword' must be a 1-D or 2-D array with last dimension between {lo} and {self._n}, not shape {codeword.shape}.")
        if codeword.ndim > 2:
            raise ValueError(f"Argument 'codeword' can be either 1-D or 2-D, not {codeword.ndim}-D.")
        return np.atleast_2d(codeword), codeword.ndim == 1

    def detect(self, codeword):
        """True where the syndrome is non-zero (_LinearCode.detect / _detect_errors, _linear.py:95-135,286-299)."""
        codeword, is_1d = self._check_codeword(codeword)
        ns = codeword.shape[-1]
        syndrome = np.matmul(codeword, self.H[:, -ns:].T.copy().view(self._field))
        detected = ~np.all(syndrome.view(np.ndarray) == 0, axis=1)
        return bool(detected[0]) if is_1d else detected

    def decode(self, codeword, erasures=None, output: str = "message", errors: bool = False):
        """_LinearCode.decode (_codes/_linear.py:137-186) → bch_decode_jit (_codes/_bch.py:1337-1578)."""
        if output not in ("message", "codeword"):
            raise ValueError(f"Argument 'output' must be in ['message', 'codeword'], not {output!r}.")
        codeword, is_1d = self._check_codeword(codeword)
        N, ns = codeword.shape
        era = None
        if erasures is not None:
            era = np.atleast_2d(np.asarray(erasures))
            if era.dtype != np.bool_:
                raise TypeError(f"Argument 'erasures' must have dtype bool, not {era.dtype}.")
            if era.shape != codeword.shape:
                raise ValueError(f"Argument 'erasures' must have shape {codeword.shape}, not {era.shape}.")
            era = np.ascontiguousarray(era).view(np.uint8)
        src = np.ascontiguousarray(codeword.view(np.ndarray), dtype=np.uint8)
        dec = _pinned.empty((N, ns), np.uint8)
        n_errors = _pinned.empty(N, np.int64)
        if N:
            st = _lib.lib().gfb_rs_decode_host(self._code(), src.ctypes.data, era.ctypes.data if era is not None else None,
                                               ns, dec.ctypes.data, n_errors.ctypes.data, N)
            _lib.check(st, "gfb_rs_decode_host")
        decoded = dec.astype(codeword.dtype, copy=False).view(self._field)
        if output == "message":
            ks = self._k - (self._n - ns)
            decoded = decoded[..., 0:ks]
        n_errors = np.array(n_errors)  # plain int64 ndarray, as in the reference
        if is_1d:
            decoded, n_errors = decoded[0, :], int(n_errors[0])
        return (decoded, n_errors) if errors else decoded

    # ------------------------------------------------------------------------------------ device-resident
    def encode_device(self, message: DeviceArray) -> DeviceArray:
        N, ks = message.shape
        out = DeviceArray.empty(self._field, (N, ks + self._n - self._k), np.uint8)
        st = _lib.lib().gfb_rs_encode(self._code(), message.data_ptr, ks, out.data_ptr, N, DeviceArray._stream())
        _lib.check(st, "gfb_rs_encode")
        return out

    def decode_device(self, codeword: DeviceArray, erasures=None):
        """Returns (decoded codewords DeviceArray [N, ns], n_errors torch.int64 tensor [N])."""
        import torch

        N, ns = codeword.shape
        out = DeviceArray.empty(self._field, (N, ns), np.uint8)
        n_errors = torch.empty(N, dtype=torch.int64, device="cuda")
        st = _lib.lib().gfb_rs_decode(self._code(), codeword.data_ptr, erasures.data_ptr() if erasures is not None else None,
                                      ns, out.data_ptr, n_errors.data_ptr(), N, DeviceArray._stream())
        _lib.check(st, "gfb_rs_decode")
        return out, n_errors
