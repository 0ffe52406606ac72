"""
galois_b200._binding — the CUDA backend as a fourth ``ufunc_mode`` of the *reference's own* field classes.

    import galois, galois_b200
    GF = galois.GF(2**31 - 1)
    galois_b200.enable(GF)            # GF.ufunc_mode == "cuda"; GF.compile("cuda") works from now on
    z = x * y                         # type(z) is GF, dtype preserved, ZeroDivisionError as in the reference
    galois.ntt(x); np.convolve(a, b); A @ B; np.linalg.inv(A); galois.ReedSolomon(255, 223).decode(c)

Nothing in /root/reference is edited: ``enable`` installs, once per process, run-time hooks on the two dispatcher
classes that ARE the reference's plugin interface:

  * ``UFunc.ufunc`` / ``UFunc.ufunc_call_only`` (_domains/_ufunc.py:69-96) select the compiled callable by
    ``field.ufunc_mode``; in mode "cuda" they return a :class:`CudaUFunc`, an object with the NumPy-ufunc call surface the
    dispatchers use (``getattr(self.ufunc, method)(*inputs, **kwargs)``, _ufunc.py:349) whose methods run
    ``gfb_binary_host`` / ``gfb_unary_host`` / ``gfb_power_host`` (include/galois_b200.h).
  * ``Function.__call__`` of the array functions on the hot path (convolve_jit, fft_jit, ifft_jit
    _domains/_function.py:111-392; matmul_jit, row_reduce_jit, lu_decompose_jit, plu_decompose_jit _domains/_linalg.py:180-428;
    evaluate_elementwise_jit, roots_jit _polys/_dense.py:404-513; bch_decode_jit _codes/_bch.py:1255-1302) is routed to
    ``gfb_convolve_host`` / ``gfb_ntt_exec_host`` / ``gfb_matmul_host`` / ``gfb_row_reduce_host`` / ``gfb_plu_decompose_host`` /
    ``gfb_poly_evaluate_host`` / ``gfb_poly_roots_host`` / ``gfb_rs_decode_host`` with native-width arrays (the reference
    round-trips through int64).  Every other ``Function`` (polynomial algebra, LFSRs, Lagrange interpolation ...) keeps
    running the reference's own JIT code, compiled against the field's previous mode (``Function.jit``, _function.py:82-95).
  * ``Array.compile`` (_domains/_array.py:322-362) accepts "cuda" for enabled classes.

``mode`` strings, ``ufunc_modes`` and ``ufunc_mode`` keep their reference meaning; ``disable(GF)`` restores the previous mode.
The CUDA side is reached through a flyweight mirror class (``galois_b200.GF`` with the same order, irreducible polynomial
and primitive element) that owns the ``gfb_field`` handle; arrays cross by ``ndarray.view`` (no copies).
"""

from __future__ import annotations

import contextlib
import ctypes
import threading

import numpy as np

from . import _lib, _pinned

_HOOK_LOCK = threading.Lock()
_hooks_installed = False
_ref: dict = {}  # reference modules / originals, filled by _install_hooks

# fields smaller than this many elements in an operation still go to the GPU by default; a maintainer can set a cutoff
# below which the reference's own compiled ufunc answers (kernel launch + PCIe latency dominate tiny arrays).
_MIN_SIZE_DEFAULT = 0


class _NoCopyArray(np.ndarray):
    """Result buffer handed back to the reference's dispatcher: ``_view_output_as_field`` (_ufunc.py:323-330) does
    ``output.astype(dtype)``, which for an identical dtype is a full extra pass over the result; skip it."""

    def astype(self, dtype, order="K", casting="unsafe", subok=True, copy=True):
        if np.dtype(dtype) == self.dtype and order in ("K", "A", "C"):
            return self.view(np.ndarray)
        return self.view(np.ndarray).astype(dtype, order=order, casting=casting, subok=subok, copy=copy)


def _nocopy(a: np.ndarray) -> np.ndarray:
    return a.view(_NoCopyArray) if isinstance(a, np.ndarray) and a.dtype != np.object_ else a


# -------------------------------------------------------------------------------------------------
# field bookkeeping
# -------------------------------------------------------------------------------------------------
def _is_reference_field(field) -> bool:
    galois = _ref.get("galois")
    return galois is not None and isinstance(field, type) and issubclass(field, galois.FieldArray) and field is not galois.FieldArray


def _mirror(field):
    """The galois_b200 field class that owns the gfb_field handle for a reference field class."""
    m = field.__dict__.get("_gfb_mirror")
    if m is None:
        from ._field import GF as MirrorGF

        if field.degree == 1:
            m = MirrorGF(int(field.order), primitive_element=int(field.primitive_element), verify=False)
        else:
            m = MirrorGF(int(field.order), irreducible_poly=int(field.irreducible_poly),
                         primitive_element=int(field.primitive_element), verify=False)
        field._gfb_mirror = m
    return m


def _check_supported(field) -> None:
    p, m = int(field.characteristic), int(field.degree)
    if m == 1:
        if field.dtypes == [np.object_]:
            raise NotImplementedError(
                f"{field.name}: the reference holds this prime field in dtype=object ('python-calculate'); "
                "the CUDA backend covers primes with integer dtypes (p < 2^31.5)."
            )
        return
    if p != 2:
        raise NotImplementedError(f"{field.name}: odd-characteristic extension fields are outside the CUDA hot path (SURVEY.md §8a).")
    if m > 128:
        raise NotImplementedError(f"{field.name}: degrees above 128 are not supported by the CUDA backend.")


def _fallback_mode(field) -> str:
    return field.__dict__.get("_gfb_prev_mode") or field._default_ufunc_mode


def _is_cuda(field) -> bool:
    return getattr(field, "_ufunc_mode", None) == "cuda"


@contextlib.contextmanager
def _reference_mode(*fields):
    """Temporarily put enabled fields back into their previous mode (used while the reference compiles or runs one of
    its own functions that has no CUDA implementation)."""
    saved = []
    for f in fields:
        if f is not None and _is_cuda(f):
            saved.append(f)
            f._ufunc_mode = _fallback_mode(f)
            if f._ufunc_mode == "jit-lookup" and f._EXP.size == 0:
                f._build_lookup_tables()
    try:
        yield
    finally:
        for f in saved:
            f._ufunc_mode = "cuda"


def enable(field, min_size: int | None = None):
    """
    Switch a reference field class (``galois.GF(...)``) to ``ufunc_mode == "cuda"``.  Returns the class.

    ``min_size``: operations on fewer elements than this are answered by the reference's own compiled ufunc (default 0:
    everything runs on the GPU).
    """
    _install_hooks()
    if not _is_reference_field(field):
        raise TypeError(f"Argument 'field' must be a galois.FieldArray subclass created by galois.GF(), not {field!r}.")
    _check_supported(field)
    _lib.lib()  # fail loudly (BackendUnavailable) without the library or a GPU: there is no CPU fallback
    _mirror(field)
    if "cuda" not in field._ufunc_modes:
        field._ufunc_modes = list(field._ufunc_modes) + ["cuda"]
    if field._ufunc_mode != "cuda":
        field._gfb_prev_mode = field._ufunc_mode or field._default_ufunc_mode
        field._ufunc_mode = "cuda"
    field._gfb_min_size = int(_MIN_SIZE_DEFAULT if min_size is None else min_size)
    return field


def disable(field):
    """Restore the mode the class had before :func:`enable`."""
    if _is_cuda(field):
        field._ufunc_mode = _fallback_mode(field)
        if field._ufunc_mode == "jit-lookup" and field._EXP.size == 0:
            field._build_lookup_tables()
    return field


def is_enabled(field) -> bool:
    return _is_cuda(field)


# -------------------------------------------------------------------------------------------------
# the ufunc object the reference's dispatchers call
# -------------------------------------------------------------------------------------------------
_BIN, _UN, _POW = "binary", "unary", "power"


class CudaUFunc:
    """
    Stands where the Numba DUFunc stands in the reference: ``UFunc.ufunc`` returns it, the dispatcher calls
    ``getattr(ufunc, method)(*ndarray_inputs, **kwargs)`` (_domains/_ufunc.py:349,403).  ``kind``/``op`` say which C-ABI
    entry point answers.  ``reference`` is the DUFunc of the field's previous mode (used for ``min_size`` and for the
    ufunc methods the GPU does not implement).
    """

    def __init__(self, dispatcher, kind: str, op: int | None):
        self.dispatcher = dispatcher
        self.field = dispatcher.field
        self.kind = kind
        self.op = op
        self.nin = 1 if kind == _UN else 2
        self.nout = 1
        self.__name__ = f"cuda_{type(dispatcher).__name__}"

    # -- helpers
    def _reference(self):
        with _reference_mode(self.field):
            return _ref["ufunc_orig"].fget(self.dispatcher)

    def _m(self):
        return _mirror(self.field)

    def _work_dtype(self, inputs):
        """Arithmetic dtype: the widest valid field dtype among the array operands (values of any valid dtype fit any
        other; the dispatcher casts the result to the dtype of the operand that owns the call, _ufunc.py:323-330)."""
        M = self._m()
        if M._wide:
            return np.dtype(object)
        valid = [np.dtype(d) for d in M.dtypes]
        best = None
        for x in inputs:
            if isinstance(x, np.ndarray) and x.dtype in valid:
                if best is None or x.dtype.itemsize > best.itemsize:
                    best = x.dtype
        return best if best is not None else valid[0]

    def _small(self, inputs) -> bool:
        ms = self.field.__dict__.get("_gfb_min_size", 0)
        if ms <= 0:
            return False
        return max((np.size(x) for x in inputs), default=1) < ms

    @staticmethod
    def _finish(result, out, where, dtype):
        res = np.asarray(result).view(np.ndarray)
        if dtype is not None and res.dtype != np.dtype(dtype):
            res = res.astype(dtype)
        if out is not None:
            target = out[0] if isinstance(out, tuple) else out
            if target is not None:
                if where is True:
                    target[...] = res
                else:
                    np.copyto(target, res, where=where, casting="unsafe")
                return target
        if where is not True:
            # elements where the mask is False are unspecified in NumPy (uninitialised memory); leave the computed value
            pass
        return _nocopy(res)

    def _as_field_operand(self, x, dtype):
        """Integer operands (scalars, ndarrays) hold field elements already (_ufunc.py:391-401 reduces them mod p)."""
        M = self._m()
        if isinstance(x, np.ndarray):
            a = x.view(np.ndarray)
            if M._wide:
                return a if a.dtype == np.object_ else a.astype(object)
            return a if a.dtype == dtype else a.astype(dtype)
        if M._wide:
            arr = np.empty((), dtype=object)
            arr[()] = int(x)
            return arr
        return np.asarray(int(x), dtype=dtype)

    # -- NumPy ufunc surface
    def __call__(self, *inputs, out=None, where=True, casting=None, dtype=None, order=None, subok=None, **_kw):
        if self._small(inputs) or where is not True:
            # where=: the reference's compiled ufuncs reject a mask ("... flagged as the ARRAYMASK ..." ValueError from
            # NumPy's iterator) while its overrides / pure-Python ufuncs honour it; hand the call to the reference's own
            # ufunc so every one of those behaviours is reproduced verbatim
            kw = {} if out is None else {"out": out}
            if where is not True:
                kw["where"] = where
            if casting is not None:
                kw["casting"] = casting
            if dtype is not None:
                kw["dtype"] = dtype
            return self._reference()(*inputs, **kw)
        M = self._m()
        wd = self._work_dtype(inputs)
        if self.kind == _UN:
            a = self._as_field_operand(inputs[0], wd)
            res = M._unary(self.op, a, wd)
        elif self.kind == _POW:
            base = self._as_field_operand(inputs[0], wd)
            expo = inputs[1]
            if isinstance(expo, np.ndarray) and expo.ndim == 0:
                expo = expo.item()
            res = M._power([base.view(M), expo], [0, 1], wd)
        else:
            a = self._as_field_operand(inputs[0], wd)
            b = self._as_field_operand(inputs[1], wd)
            res = M._binary(self.op, a, b, wd)
        return self._finish(res, out, where, dtype)

    def outer(self, a, b, **kwargs):
        a = np.asarray(a)
        b = np.asarray(b)
        return self(a.reshape(a.shape + (1,) * b.ndim), b, **kwargs)

    def reduce(self, a, axis=0, dtype=None, out=None, keepdims=False, **kwargs):
        if self.kind != _BIN or kwargs.get("where", True) is not True or "initial" in kwargs or self._small((a,)):
            return self._reference().reduce(a, axis=axis, dtype=dtype, out=out, keepdims=keepdims, **kwargs)
        M = self._m()
        wd = self._work_dtype((a,))
        arr = self._as_field_operand(a, wd)
        if isinstance(axis, tuple):
            res = arr.view(M)
            for ax in sorted((x % arr.ndim for x in axis), reverse=True):
                res = M._reduce(self.op, np.asarray(res).view(M), ax, keepdims, wd)
        else:
            res = M._reduce(self.op, arr.view(M), axis, keepdims, wd)
        res = np.asarray(res).view(np.ndarray)
        if out is not None:
            target = out[0] if isinstance(out, tuple) else out
            target[...] = res
            return target
        return res if res.ndim else res[()]

    def accumulate(self, a, axis=0, dtype=None, out=None, **kwargs):
        M = self._m()
        if self.kind != _BIN or self._small((a,)) or not hasattr(M, "_accumulate"):
            return self._reference().accumulate(a, axis=axis, dtype=dtype, out=out, **kwargs)
        wd = self._work_dtype((a,))
        res = np.asarray(M._accumulate(self.op, self._as_field_operand(a, wd).view(M), axis, wd)).view(np.ndarray)
        return self._finish(res, out, True, dtype)

    def reduceat(self, a, indices, axis=0, dtype=None, out=None, **kwargs):
        M = self._m()
        if self.kind != _BIN or self._small((a,)) or not hasattr(M, "_reduceat"):
            return self._reference().reduceat(a, indices, axis=axis, dtype=dtype, out=out, **kwargs)
        wd = self._work_dtype((a,))
        res = np.asarray(M._reduceat(self.op, self._as_field_operand(a, wd).view(M), indices, axis, wd)).view(np.ndarray)
        return self._finish(res, out, True, dtype)

    def at(self, a, indices, b=None):
        M = self._m()
        if self._small((a,)) or not hasattr(M, "_at"):
            args = (a, indices) if b is None else (a, indices, b)
            return self._reference().at(*args)
        wd = self._work_dtype((a,))
        if not (isinstance(a, np.ndarray) and a.dtype == wd):
            args = (a, indices) if b is None else (a, indices, b)
            return self._reference().at(*args)
        if b is not None and self.kind == _BIN:
            b = self._as_field_operand(b, wd).view(M)
        M._at(self.kind, self.op, a.view(M), indices, b)
        return None


def _classify(dispatcher):
    """(kind, op) of a reference ufunc dispatcher, or None when the GPU has no kernel for it (log, sqrt ...)."""
    u = _ref["ufunc_mod"]
    if isinstance(dispatcher, u.add_ufunc):
        return _BIN, _lib.GFB_ADD
    if isinstance(dispatcher, u.subtract_ufunc):
        return _BIN, _lib.GFB_SUB
    if isinstance(dispatcher, u.multiply_ufunc):
        return _BIN, _lib.GFB_MUL
    if isinstance(dispatcher, u.divide_ufunc):
        return _BIN, _lib.GFB_DIV
    if isinstance(dispatcher, u.negative_ufunc):
        return _UN, _lib.GFB_NEG
    if isinstance(dispatcher, u.reciprocal_ufunc):
        return _UN, _lib.GFB_INV
    if isinstance(dispatcher, u.power_ufunc):
        return _POW, None
    return None


def _cuda_ufunc(dispatcher):
    cu = dispatcher.__dict__.get("_gfb_cuda")
    if cu is None:
        cls = _classify(dispatcher)
        cu = CudaUFunc(dispatcher, *cls) if cls is not None else False
        dispatcher._gfb_cuda = cu
    return cu or None


# -------------------------------------------------------------------------------------------------
# array functions (Function.__call__ level)
# -------------------------------------------------------------------------------------------------
def _to_m(M, x):
    return np.asarray(x).view(np.ndarray).view(M)


def _to_ref(field, x):
    return field._view(np.asarray(x).view(np.ndarray))


def _cuda_convolve(self, a, b, mode="full"):
    from . import _linalg

    vi = _ref["verify_isinstance"]
    vi(a, self.field)
    vi(b, self.field)
    if not mode == "full":
        raise ValueError(f"Operation 'convolve' currently only supports mode of 'full', not {mode!r}.")
    if a.ndim != 1 or b.ndim != 1 or a.size == 0 or b.size == 0:
        return NotImplemented
    M = _mirror(self.field)
    return _to_ref(self.field, _linalg.convolve(M, _to_m(M, a), _to_m(M, b)))


def _cuda_fft(self, x, n=None, axis=-1, norm=None):
    from . import _ntt

    _ref["verify_isinstance"](x, self.field)
    M = _mirror(self.field)
    if M._wide or x.ndim != 1:
        return NotImplemented
    size = x.size if n is None else n
    if size < 1 or (int(self.field.order) - 1) % size != 0:
        self.field.primitive_root_of_unity(size)  # raises the reference's own error
    y = _ntt._fft_dispatch(M, self._direction == "backward", _to_m(M, x), n=n, axis=axis, norm=norm)
    return _to_ref(self.field, y)


def _cuda_matmul(self, A, B, out=None, **kwargs):
    from . import _linalg

    vi = _ref["verify_isinstance"]
    vi(A, self.field)
    vi(B, self.field)
    if A.ndim < 1 or B.ndim < 1 or A.shape[-1] != (B.shape[-2] if B.ndim > 1 else B.shape[0]):
        return NotImplemented  # let the reference raise its own error (NumPy's gufunc text for prime fields, _linalg.py:41-67)
    try:
        np.broadcast_shapes(A.shape[:-2], B.shape[:-2])
    except ValueError:
        return NotImplemented
    M = _mirror(self.field)
    C = _linalg.matmul(M, _to_m(M, A), _to_m(M, B))
    C = _to_ref(self.field, C) if isinstance(C, np.ndarray) and C.ndim else self.field(int(C), dtype=A.dtype)
    if out is not None:
        assert isinstance(out, tuple) and len(out) == 1
        out[0][...] = C[...]
        return out[0]
    return C


def _cuda_row_reduce(self, A, ncols=None):
    from . import _linalg

    _ref["verify_isinstance"](A, self.field)
    M = _mirror(self.field)
    if M._wide:
        return NotImplemented
    R, p = _linalg._row_reduce_raw(M, _to_m(M, A) if A.ndim == 2 else A, ncols)
    return _to_ref(self.field, R), p


def _cuda_lu(self, A):
    from . import _linalg

    _ref["verify_isinstance"](A, self.field)
    M = _mirror(self.field)
    if M._wide or A.ndim != 2 or A.dtype != np.dtype(self.field.dtypes[0]):
        # the reference builds L with the field's default dtype (Identity(m), _linalg.py:367); keep its mixed-dtype result
        return NotImplemented
    L, U, _, _ = _linalg._plu_raw(M, _to_m(M, A), pivoting=False)
    return _to_ref(self.field, L), _to_ref(self.field, U)


def _cuda_plu(self, A):
    from . import _linalg

    _ref["verify_isinstance"](A, self.field)
    M = _mirror(self.field)
    if M._wide or A.ndim != 2 or A.dtype != np.dtype(self.field.dtypes[0]) or 0 in A.shape:
        return NotImplemented
    L, U, perm, n_swaps = _linalg._plu_raw(M, _to_m(M, A), pivoting=True)
    m = A.shape[0]
    P_rows = np.zeros((m, m), dtype=A.dtype)
    P_rows[np.arange(m), perm] = 1
    return _to_ref(self.field, P_rows).T, _to_ref(self.field, L), _to_ref(self.field, U), n_swaps


def _cuda_poly_evaluate(self, coeffs, x):
    from . import _linalg

    M = _mirror(self.field)
    if not isinstance(coeffs, self.field) or not isinstance(x, self.field) or coeffs.ndim != 1 or coeffs.size == 0:
        return NotImplemented
    y = _linalg.poly_evaluate(_to_m(M, coeffs), _to_m(M, x))
    y = _to_ref(self.field, np.asarray(y).reshape(x.shape))
    return y if y.dtype == x.dtype else y.astype(x.dtype)


def _cuda_poly_roots(self, nonzero_degrees, nonzero_coeffs):
    from . import _poly

    vi = _ref["verify_isinstance"]
    vi(nonzero_degrees, np.ndarray)
    vi(nonzero_coeffs, self.field)
    M = _mirror(self.field)
    if M._wide or self.field.order > 2**33 or nonzero_degrees.size == 0:
        return NotImplemented
    return _to_ref(self.field, _poly.poly_roots(np.asarray(nonzero_degrees), _to_m(M, nonzero_coeffs)))


_CODES: dict = {}
_CODES_LOCK = threading.Lock()


def _decode_code(ext_field, base_is_gf2: bool, n: int, alpha: int, c: int, n_roots: int):
    """gfb_rs_code* for bch_decode_jit's (design_n, alpha, c, roots) over ``ext_field`` (decode-only handle)."""
    lib = _lib.lib()
    M = _mirror(ext_field)
    h = M._handle()
    key = (h, base_is_gf2, n, alpha, c, n_roots)
    with _CODES_LOCK:
        code = _CODES.get(key)
        if code is None:
            out = ctypes.c_void_p()
            d = n_roots + 1
            k = n - n_roots
            if base_is_gf2:
                gen = np.zeros(d, dtype=np.uint8)  # the decoder never touches g(x): x^(n-k) + 1 keeps the constructor happy
                gen[0] = gen[-1] = 1
                st = lib.gfb_bch_create(ctypes.byref(out), h, n, k, d, c, ctypes.c_uint64(alpha), gen.ctypes.data)
            else:
                st = lib.gfb_rs_create(ctypes.byref(out), h, n, k, c, ctypes.c_uint64(alpha))
            _lib.check(st, f"gfb_rs_create({n},{k})")
            code = out.value
            _CODES[key] = code
    return code


def _cuda_bch_decode(self, codeword, erasures, design_n, alpha, c, roots):
    field, ext = self.field, self.extension_field
    base_is_gf2 = field is not ext
    if base_is_gf2 and int(field.order) != 2:
        return NotImplemented  # non-binary BCH: symbols in GF(q), q > 2, syndromes in GF(q^m)
    if int(ext.characteristic) != 2 or int(ext.degree) > 8 or codeword.ndim != 2 or len(roots) < 1:
        return NotImplemented
    N, ns = codeword.shape
    n_roots = int(len(roots))
    if not n_roots < ns <= design_n:
        return NotImplemented
    code = _decode_code(ext, base_is_gf2, int(design_n), int(alpha), int(c), n_roots)
    src = np.ascontiguousarray(codeword.view(np.ndarray), dtype=np.uint8)
    era = None
    if erasures is not None:
        era = np.ascontiguousarray(np.asarray(erasures).astype(np.bool_, copy=False)).view(np.uint8)
        if not era.any():
            era = None
    dec = _pinned.empty((N, ns), np.uint8)
    n_errors = np.empty(N, dtype=np.int64)
    if N:
        st = _lib.lib().gfb_rs_decode_host(code, src.ctypes.data, era.ctypes.data if era is not None else None, ns,
                                           dec.ctypes.data, n_errors.ctypes.data, N)
        _lib.check(st, "gfb_rs_decode_host")
    dec = np.array(dec, dtype=codeword.dtype)
    return dec.view(field), n_errors  # .view(field) validates GF(2) symbols exactly as _bch.py:1300 does


def _function_fields(fn):
    fields = [getattr(fn, "field", None)]
    ext = getattr(fn, "extension_field", None)
    if ext is not None and ext is not fields[0]:
        fields.append(ext)
    return [f for f in fields if f is not None]


def _wrap_function_call(cls, cuda_impl):
    orig = cls.__dict__.get("__call__")
    if orig is None or getattr(orig, "_gfb_wrapped", False):
        return

    def __call__(self, *args, **kwargs):
        fields = _function_fields(self)
        mode_field = fields[-1]  # bch_decode_jit branches on the extension field's mode (_bch.py:1280)
        if not _is_cuda(mode_field):
            if any(_is_cuda(f) for f in fields):
                with _reference_mode(*fields):
                    return orig(self, *args, **kwargs)
            return orig(self, *args, **kwargs)
        if cuda_impl is not None:
            r = cuda_impl(self, *args, **kwargs)
            if r is not NotImplemented:
                return r
        if any(_fallback_mode(f) == "python-calculate" for f in fields):
            # object-dtype fields: the reference's pure-Python implementation (its __call__ branches on the mode)
            with _reference_mode(*fields):
                return orig(self, *args, **kwargs)
        return orig(self, *args, **kwargs)

    __call__._gfb_wrapped = True
    __call__.__wrapped__ = orig
    __call__.__doc__ = orig.__doc__
    cls.__call__ = __call__


def _all_subclasses(cls):
    seen, stack = [], [cls]
    while stack:
        c = stack.pop()
        for s in c.__subclasses__():
            if s not in seen:
                seen.append(s)
                stack.append(s)
    return seen


# -------------------------------------------------------------------------------------------------
# hook installation
# -------------------------------------------------------------------------------------------------
def _install_hooks():
    global _hooks_installed
    with _HOOK_LOCK:
        if _hooks_installed:
            return
        import galois
        from galois._codes import _bch as r_bch
        from galois._domains import _array as r_array
        from galois._domains import _function as r_function
        from galois._domains import _linalg as r_linalg
        from galois._domains import _ufunc as r_ufunc
        from galois._helper import verify_isinstance
        from galois._polys import _dense as r_dense

        _ref.update(galois=galois, ufunc_mod=r_ufunc, verify_isinstance=verify_isinstance)
        UFunc, Function, Array = r_ufunc.UFunc, r_function.Function, r_array.Array

        # 1. ufunc selection (_domains/_ufunc.py:69-96)
        ufunc_orig, call_only_orig = UFunc.__dict__["ufunc"], UFunc.__dict__["ufunc_call_only"]
        _ref.update(ufunc_orig=ufunc_orig, call_only_orig=call_only_orig)

        def _select(self, orig_prop):
            if _is_cuda(self.field):
                if self.override:
                    return self.override  # np.bitwise_xor etc.: the reference's own override in every mode
                cu = _cuda_ufunc(self)
                if cu is not None:
                    return cu
                with _reference_mode(self.field):
                    return orig_prop.fget(self)
            return orig_prop.fget(self)

        UFunc.ufunc = property(lambda self: _select(self, ufunc_orig), doc=ufunc_orig.__doc__)
        UFunc.ufunc_call_only = property(lambda self: _select(self, call_only_orig), doc=call_only_orig.__doc__)

        # 2. fused division: the reference multiplies by a separately computed reciprocal (_ufunc.py:430-441); in mode
        #    "cuda" the divide kernel does both in one pass over the operands.
        div_orig = r_ufunc.divide_ufunc.__call__

        def divide_call(self, ufunc, method, inputs, kwargs, meta):
            if _is_cuda(self.field) and method == "__call__" and not self.override and kwargs.get("where", True) is True:
                self._verify_operands_in_same_field(ufunc, inputs, meta)
                inputs, kwargs = self._view_inputs_as_ndarray(inputs, kwargs)
                output = self.ufunc(*inputs, **kwargs)
                return self._view_output_as_field(output, self.field, meta["dtype"])
            return div_orig(self, ufunc, method, inputs, kwargs, meta)

        r_ufunc.divide_ufunc.__call__ = divide_call

        # 3. compiled array functions: the reference's own JIT code for everything without a CUDA implementation
        jit_orig, function_orig = Function.__dict__["jit"], Function.__dict__["function"]

        def jit(self):
            fields = _function_fields(self)
            if any(_is_cuda(f) for f in fields):
                with _reference_mode(*fields):
                    return jit_orig.fget(self)
            return jit_orig.fget(self)

        def function(self):
            fields = _function_fields(self)
            if any(_is_cuda(f) for f in fields):
                with _reference_mode(*fields):
                    return function_orig.fget(self)
            return function_orig.fget(self)

        Function.jit = property(jit, doc=jit_orig.__doc__)
        Function.function = property(function, doc=function_orig.__doc__)

        cuda_calls = {
            r_function.convolve_jit: _cuda_convolve,
            r_function.fft_jit: _cuda_fft,
            r_function.ifft_jit: _cuda_fft,
            r_linalg.matmul_jit: _cuda_matmul,
            r_linalg.row_reduce_jit: _cuda_row_reduce,
            r_linalg.lu_decompose_jit: _cuda_lu,
            r_linalg.plu_decompose_jit: _cuda_plu,
            r_dense.evaluate_elementwise_jit: _cuda_poly_evaluate,
            r_dense.roots_jit: _cuda_poly_roots,
            r_bch.bch_decode_jit: _cuda_bch_decode,
        }
        _ref["cuda_calls"] = cuda_calls
        for sub in _all_subclasses(Function):
            _wrap_function_call(sub, cuda_calls.get(sub))

        # 4. Array.compile (_domains/_array.py:322-362) learns the new mode
        compile_orig = Array.__dict__["compile"].__func__

        def compile(cls, mode):  # noqa: A001
            if mode == "cuda":
                if "cuda" not in cls._ufunc_modes:
                    raise ValueError(
                        f"Argument 'mode' must be in {cls.ufunc_modes} for {cls._name}, not 'cuda'. "
                        "Call galois_b200.enable(field) first."
                    )
                enable(cls, cls.__dict__.get("_gfb_min_size"))
                return None
            if _is_cuda(cls):
                cls._ufunc_mode = _fallback_mode(cls)
            return compile_orig(cls, mode)

        compile.__doc__ = compile_orig.__doc__
        Array.compile = classmethod(compile)

        _hooks_installed = True
