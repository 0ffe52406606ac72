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
        """True where the word is not a codeword (_LinearCode.detect, _linear.py:95-135).  The reference multiplies by
        H; for a systematic code "syndrome != 0" is the same predicate as "re-encoding the message part changes the
        parity part", which reuses the encode kernel."""
        codeword, is_1d = self._check_codeword(codeword)
        ns = codeword.shape[-1]
        ks = self._k - (self._n - ns)
        again = self.encode(codeword[:, :ks])
        detected = np.any(np.asarray(again) != np.asarray(codeword), axis=1)
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
        self._validate_decoded(dec)
        decoded = dec.astype(codeword.dtype, copy=False).view(self._field)
        if output == "message":
            ks = self._k - (self._n - ns)
            decoded = decoded[..., 0:ks]
        n_errors = np.array(n_errors)  # plain int64 ndarray, as in the reference
        if is_1d:
            decoded, n_errors = decoded[0, :], int(n_errors[0])
        return (decoded, n_errors) if errors else decoded

    def _validate_decoded(self, dec: np.ndarray) -> None:
        """Hook for codes whose symbol field is smaller than the decoding field (BCH)."""

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


# -------------------------------------------------------------------------------------------------
# Binary BCH codes — SURVEY.md §8(f) rank 1: the same decoder kernel with field != extension_field
# -------------------------------------------------------------------------------------------------
def _minimal_poly_gf2(beta: int, ext_poly: int) -> int:
    """Minimal polynomial over GF(2) of beta in GF(2^m) = GF(2)[x]/(ext_poly), as a GF(2)[x] bit mask
    (FieldArray.minimal_poly; product of (x - conjugate) over the Frobenius orbit of beta)."""
    conj, b = [], beta
    while b not in conj:
        conj.append(b)
        b = _nt.mulmod2(b, b, ext_poly)
    coeffs = [1]  # ascending, coefficients in GF(2^m)
    for cj in conj:
        nxt = [0] * (len(coeffs) + 1)
        for i, a in enumerate(coeffs):
            nxt[i + 1] ^= a  # a * x
            nxt[i] ^= _nt.mulmod2(a, cj, ext_poly)  # a * conj  (minus = plus in characteristic 2)
        coeffs = nxt
    if any(v > 1 for v in coeffs):
        raise ArithmeticError("minimal polynomial has a coefficient outside GF(2)")
    return sum(v << i for i, v in enumerate(coeffs))


def _generator_poly_from_d(d: int, alpha: int, c: int, ext_poly: int) -> tuple[int, list[int]]:
    """_codes/_bch.py:1178-1197: g(x) = product of the distinct minimal polynomials of alpha^c .. alpha^(c+d-2)."""
    g, roots, seen = 1, [], []
    for i in range(d - 1):
        root = _nt.powmod2(alpha, c + i, ext_poly)
        roots.append(root)
        mi = _minimal_poly_gf2(root, ext_poly)
        if mi not in seen:
            g = _nt.clmul(g, mi)
            seen.append(mi)
    return g, roots


def _generator_poly_from_k(n: int, k: int, m: int, alpha: int, c: int, ext_poly: int) -> tuple[int, list[int]]:
    """_codes/_bch.py:1200-1252: binary search for a design distance giving n - k parity symbols, then the largest."""
    deg = lambda g: g.bit_length() - 1  # noqa: E731
    possible = list(range((n - k) // m + 1, (n - k) + 2))
    g = roots = d = None
    while possible:
        idx = len(possible) // 2
        d = possible[idx]
        g, roots = _generator_poly_from_d(d, alpha, c, ext_poly)
        if deg(g) < n - k:
            possible = possible[idx + 1:]
        elif deg(g) == n - k:
            break
        else:
            possible = possible[:idx]
    else:
        raise ValueError(f"The BCH({n}, {k}) code over GF(2) with alpha={alpha} and c={c} does not exist.")
    best = (g, roots)
    while True:
        d += 1
        g, roots = _generator_poly_from_d(d, alpha, c, ext_poly)
        if deg(g) == n - k:
            best = (g, roots)
        elif deg(g) > n - k:
            break
    return best


class BCH(ReedSolomon):
    """
    A binary BCH(n, k) code: symbols in GF(2), syndromes in GF(2^m), m <= 8.  Mirrors galois.BCH
    (_codes/_bch.py:27-240) for ``field = GF(2)``; encode / decode run the same CUDA kernels as ReedSolomon.
    """

    def __init__(self, n: int, k: int | None = None, d: int | None = None, field=None, extension_field=None,
                 alpha=None, c: int = 1, systematic: bool = True):
        for name, v, opt in (("n", n, False), ("k", k, True), ("d", d, True), ("c", c, False)):
            if not (isinstance(v, (int, np.integer)) or (opt and v is None)):
                raise TypeError(f"Argument {name!r} must be an int, not {type(v)}.")
        if d is not None and not d >= 1:
            raise ValueError(f"Argument 'd' must be at least 1, not {d}.")
        if not c >= 0:
            raise ValueError(f"Argument 'c' must be at least 0, not {c}.")
        if not systematic:
            raise NotImplementedError("Non-systematic BCH codes are outside the CUDA hot path.")
        if field is None:
            field = GF(2)
        if not field._is_prime_field:
            raise ValueError("Current BCH codes over GF(q) for prime power q are not supported. Proper Galois field towers are needed first.")
        if field.order != 2:
            raise NotImplementedError("Only binary BCH codes (field = GF(2)) are built on the CUDA path (SURVEY.md §8f).")
        if extension_field is None:
            m = _nt.ilog(n, 2) + 1
            extension_field = GF(2**m, irreducible_poly=_MATLAB_PRIMITIVE_POLY_2[m])
        if extension_field.characteristic != 2 or extension_field.degree > 8:
            raise NotImplementedError("The CUDA BCH kernels cover extension fields GF(2^m), m <= 8.")
        if (extension_field.order - 1) % n != 0:
            raise ValueError(f"Argument 'n' must divide q^m - 1 = {extension_field.order - 1}, not {n}.")
        alpha = extension_field.primitive_root_of_unity(n) if alpha is None else extension_field(alpha)
        ext_poly = int(extension_field.irreducible_poly)
        if d is not None:
            g, roots = _generator_poly_from_d(d, int(alpha), c, ext_poly)
            kk = n - (g.bit_length() - 1)
            if k not in (None, kk):
                raise ValueError(f"The requested [{n}, {k}, {d}] code is not consistent. When designing the code with design distance {d}, the resulting code is [{n}, {kk}, {d}].")
            k = kk
        elif k is not None:
            g, roots = _generator_poly_from_k(n, k, extension_field.degree, int(alpha), c, ext_poly)
            d = len(roots) + 1
        else:
            raise ValueError("Argument 'k' or 'd' must be provided to define the code size.")
        self._field, self._extension_field = field, extension_field
        self._n, self._k, self._d = int(n), int(k), int(d)
        self._alpha, self._c, self._systematic = alpha, int(c), True
        self._is_primitive = n == extension_field.order - 1
        self._is_narrow_sense = c == 1
        self._roots = extension_field(roots) if roots else extension_field([])
        self._g_int = g
        self._codes = {}
        self._generator_poly = None
        self._P = None

    extension_field = property(lambda self: self._extension_field)

    def _code(self):
        lib = _lib.lib()
        h = self._extension_field._handle()
        code = self._codes.get(h)
        if code is None:
            r = self._n - self._k
            gen = np.array([(self._g_int >> (r - i)) & 1 for i in range(r + 1)], dtype=np.uint8)  # degree-descending
            out = ctypes.c_void_p()
            st = lib.gfb_bch_create(ctypes.byref(out), h, self._n, self._k, self._d, self._c,
                                    ctypes.c_uint64(int(self._alpha)), gen.ctypes.data)
            _lib.check(st, f"gfb_bch_create({self._n},{self._k},{self._d})")
            code = out.value
            self._codes[h] = code
        return code

    @property
    def roots(self) -> FieldArray:
        return self._roots

    @property
    def generator_poly(self) -> FieldArray:
        r = self._n - self._k
        return self._field([(self._g_int >> (r - i)) & 1 for i in range(r + 1)])

    @property
    def H(self):
        raise NotImplementedError("BCH.H (parity-check matrix from h(x)) is host-side symbolic work outside the CUDA path.")

    def _validate_decoded(self, dec: np.ndarray) -> None:
        # Beyond the design distance the decoder can "correct" to a word with symbols outside GF(2); the reference
        # then raises from `dec_codeword.view(self.field)` (_bch.py:1300 -> _fields/_array.py:177).  Same here.
        if dec.size and int(dec.max()) >= 2:
            bad = dec[dec >= 2]
            raise ValueError(f"GF(2) arrays must have elements in `0 <= x < 2`, not {bad}.")

    def __repr__(self):
        return f"<BCH Code: [{self._n}, {self._k}, {self._d}] over GF(2)>"
