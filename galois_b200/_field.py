"""
Host-side mirror of the reference's field-array interface, backed by the CUDA C ABI.

``GF = galois_b200.GF(p**m)`` returns a ``FieldArray`` subclass — an ``np.ndarray`` subclass like the
reference's (/root/reference/src/galois/_fields/_array.py, _domains/_array.py) — whose NumPy ufuncs are
dispatched the way ``UFuncMixin.__array_ufunc__`` does (_domains/_ufunc.py:660-713), but the compiled
callable behind each dispatcher is a hand-written sm_100a kernel reached through
``include/galois_b200.h`` instead of a Numba DUFunc.  Same names, argument meaning, result dtype
rules and exceptions; no CPU fallback (``BackendUnavailable`` if the library or the GPU is missing).

Two residencies:
  * ``FieldArray``  — host ndarray (the literal drop-in); every op runs the `*_host` pipeline
    (H2D → kernel → D2H).
  * ``DeviceArray`` — device-resident handle (``GF.device(x)`` / ``x.to_device()``) for chains of ops
    without PCIe traffic; ``.cpu()`` returns a ``FieldArray``.
"""

from __future__ import annotations

import ctypes
import threading
import weakref

import numpy as np

from . import _lib, _nt, _pinned

_DTYPES = [np.uint8, np.uint16, np.uint32, np.int8, np.int16, np.int32, np.int64]  # _domains/_meta.py:19

_OPS_BINARY = {np.add: _lib.GFB_ADD, np.subtract: _lib.GFB_SUB, np.multiply: _lib.GFB_MUL,
               np.true_divide: _lib.GFB_DIV, np.floor_divide: _lib.GFB_DIV}
_OPS_UNARY = {np.negative: _lib.GFB_NEG, np.reciprocal: _lib.GFB_INV, np.square: _lib.GFB_SQUARE}

# ufuncs the reference refuses on field arrays (_domains/_ufunc.py:565-605)
_UNSUPPORTED_UFUNCS = [
    np.float_power, np.fmod, np.modf, np.fmin, np.fmax, np.exp, np.exp2, np.expm1, np.log2, np.log10, np.log1p,
    np.cbrt, np.logaddexp, np.logaddexp2, np.sin, np.cos, np.tan, np.arcsin, np.arccos, np.arctan, np.sinh, np.cosh,
    np.tanh, np.heaviside, np.gcd, np.lcm, np.rint, np.sign, np.conj, np.conjugate, np.absolute, np.fabs, np.floor,
    np.ceil, np.trunc,
]
_UFUNCS_REQUIRING_VIEW = [np.bitwise_and, np.bitwise_or, np.bitwise_xor, np.left_shift, np.right_shift, np.positive]


class FieldMeta(type):
    """Class-level properties and repr of a field (the reference's FieldArrayMeta, _fields/_meta.py)."""

    def __repr__(cls):
        return f"<class 'galois_b200.{cls.name}'>" if getattr(cls, "order", None) else super().__repr__()

    def __str__(cls):
        return repr(cls)

    @property
    def properties(cls) -> str:
        return (
            f"Galois Field:\n  name: {cls.name}\n  characteristic: {cls.characteristic}\n  degree: {cls.degree}\n"
            f"  order: {cls.order}\n  irreducible_poly: {cls.irreducible_poly}\n  is_primitive_poly: n/a\n"
            f"  primitive_element: {cls.primitive_element}\n  ufunc_mode: {cls.ufunc_mode}"
        )


def _as_u64(v: int) -> ctypes.c_uint64:
    return ctypes.c_uint64(v & 0xFFFFFFFFFFFFFFFF)


# np functions answered by galois_b200._linalg (LinalgFunctionMixin._OVERRIDDEN_FUNCTIONS, _domains/_linalg.py:562-575)
_LINALG_FUNCTIONS = {
    np.dot: "dot",
    np.vdot: "vdot",
    np.inner: "inner",
    np.outer: "outer",
    np.linalg.det: "det",
    np.linalg.matrix_rank: "matrix_rank",
    np.linalg.solve: "solve",
    np.linalg.inv: "inv",
}


class FieldArray(np.ndarray, metaclass=FieldMeta):
    """An array over GF(p^m).  Construct concrete classes with :func:`galois_b200.GF`."""

    # class attributes filled in by the factory ------------------------------------------------
    name = ""
    characteristic = 0
    degree = 0
    order = 0
    irreducible_poly = 0  # integer representation
    primitive_element = 0
    dtypes: list = []
    ufunc_mode = "cuda"
    ufunc_modes = ["cuda"]
    _is_prime_field = False
    _wide = False  # 64 <= m <= 128: dtype=object on the host, two uint64 limbs on the device
    _handles: dict = {}  # device index -> gfb_field*
    _lock = threading.Lock()

    # ------------------------------------------------------------------------------ construction
    def __new__(cls, x, dtype=None, copy=True, order="K", ndmin=0):
        if cls is FieldArray:
            raise NotImplementedError("FieldArray is abstract: create a field class with galois_b200.GF(order).")
        dtype = cls._get_dtype(dtype)
        x = cls._verify_array_like(x, dtype)
        arr = np.array(x, dtype=dtype, copy=copy, order=order, ndmin=ndmin)
        return arr.view(cls)

    @classmethod
    def _get_dtype(cls, dtype):
        if dtype is None:
            return np.dtype(cls.dtypes[0])
        dtype = np.dtype(dtype)
        if dtype not in [np.dtype(d) for d in cls.dtypes]:
            raise TypeError(f"{cls.name} arrays only support dtypes {[np.dtype(d).name for d in cls.dtypes]}, not {dtype.name!r}.")
        return dtype

    @classmethod
    def _verify_array_like(cls, x, dtype):
        if isinstance(x, cls):
            return x.view(np.ndarray)
        if isinstance(x, (int, np.integer)):
            cls._verify_scalar(int(x))
            return x
        if isinstance(x, (list, tuple)):
            x = np.array(x, dtype=object if cls._wide else None)
            if x.size == 0:
                x = x.astype(dtype)  # np.array([]) is float64; an empty field array is fine
        if isinstance(x, np.ndarray):
            if x.dtype == np.object_:
                flat = [int(v) for v in x.ravel()]
                if any(v < 0 or v >= cls.order for v in flat):
                    raise ValueError(f"{cls.name} arrays must have elements in `0 <= x < {cls.order}`.")
                return x if cls._wide else x.astype(np.int64)
            if not np.issubdtype(x.dtype, np.integer):
                raise TypeError(f"{cls.name} arrays must have integer dtypes, not {x.dtype}.")
            if x.size and (int(x.min()) < 0 or int(x.max()) >= cls.order):
                raise ValueError(f"{cls.name} arrays must have elements in `0 <= x < {cls.order}`.")
            return x
        raise TypeError(f"{cls.name} arrays can be created with scalars of type int, lists/tuples, or ndarrays, not {type(x)}.")

    @classmethod
    def _verify_scalar(cls, v: int):
        if not 0 <= v < cls.order:
            raise ValueError(f"{cls.name} scalars must be in `0 <= x < {cls.order}`, not {v}.")

    @classmethod
    def _view(cls, array: np.ndarray) -> "FieldArray":
        return array.view(cls)

    def __array_finalize__(self, obj):
        if obj is not None and not isinstance(obj, FieldArray) and type(self) is not FieldArray:
            # a plain ndarray being viewed as a field: the reference checks dtype here (_domains/_array.py:108-134)
            if self.dtype not in [np.dtype(d) for d in type(self).dtypes]:
                raise TypeError(f"{type(self).name} can only have integer dtypes {[np.dtype(d).name for d in type(self).dtypes]}, not {self.dtype}.")

    # getters / setters / casts (_domains/_array.py:413-451) --------------------------------------
    def __getitem__(self, key):
        """A single element comes back as a 0-d field array, not a bare NumPy integer, so arithmetic on it stays in
        the field (_domains/_array.py:413-421)."""
        item = super().__getitem__(key)
        if np.isscalar(item):
            item = type(self)(item, dtype=self.dtype)
        return item

    def __setitem__(self, key, value):
        """Assigned values must be field elements: the kernels rely on 0 <= x < order (_domains/_array.py:423-429)."""
        cls = type(self)
        if isinstance(value, FieldArray):
            if not isinstance(value, cls):
                raise TypeError(f"{cls.name} arrays can only be assigned elements of {cls.name}, not {type(value)}.")
        else:
            value = cls._verify_array_like(value, self.dtype)
        super().__setitem__(key, value)

    def astype(self, dtype, order="K", casting="unsafe", subok=True, copy=True):
        """Only the field's own integer dtypes are valid targets (_domains/_array.py:445-451)."""
        cls = type(self)
        if np.dtype(dtype) not in [np.dtype(d) for d in cls.dtypes]:
            raise TypeError(
                f"{cls.name} arrays can only be cast as integer dtypes in {[np.dtype(d).name for d in cls.dtypes]}, not {np.dtype(dtype).name}."
            )
        return super().astype(dtype, order=order, casting=casting, subok=subok, copy=copy)

    # alternate constructors (_domains/_array.py:188-320) ----------------------------------------
    @classmethod
    def Zeros(cls, shape, dtype=None):
        dtype = cls._get_dtype(dtype)
        a = _pinned.empty(shape, dtype)
        a[...] = 0
        return a.view(cls)

    @classmethod
    def Ones(cls, shape, dtype=None):
        dtype = cls._get_dtype(dtype)
        a = _pinned.empty(shape, dtype)
        a[...] = 1
        return a.view(cls)

    @classmethod
    def Range(cls, start, stop, step=1, dtype=None):
        dtype = cls._get_dtype(dtype)
        if not 0 <= start <= cls.order or not 0 <= stop <= cls.order:
            raise ValueError(f"Arguments must be within the field's order {cls.order}.")
        return np.arange(start, stop, step=step, dtype=dtype).view(cls)

    @classmethod
    def Random(cls, shape=(), low=0, high=None, seed=None, dtype=None):
        """Same generator and stream as the reference: default_rng(seed).integers(low, high, dtype=dtype)
        (_domains/_array.py:243-283)."""
        dtype = cls._get_dtype(dtype)
        high = cls.order if high is None else high
        if not 0 <= low < high <= cls.order:
            raise ValueError(f"Arguments must satisfy `0 <= low < high <= order`, not `0 <= {low} < {high} <= {cls.order}`.")
        rng = np.random.default_rng(seed)
        if dtype != np.object_:
            vals = rng.integers(low, high, shape, dtype=dtype)
            out = _pinned.empty(vals.shape, dtype)
            out[...] = vals
            return out.view(cls)
        # object dtype: the reference draws uniformly on Python ints (_domains/_array.py:281-283)
        import random as _random

        rnd = _random.Random(int(rng.integers(0, 2**63)))
        shape_t = (shape,) if isinstance(shape, int) else tuple(shape)
        n = int(np.prod(shape_t)) if shape_t else 1
        arr = np.empty(n, dtype=object)
        for i in range(n):
            arr[i] = rnd.randrange(low, high)
        return arr.reshape(shape_t).view(cls)

    @classmethod
    def Identity(cls, size, dtype=None):
        dtype = cls._get_dtype(dtype)
        return np.identity(size, dtype=dtype).view(cls)

    # ------------------------------------------------------------------------------ field facts
    @classmethod
    def primitive_root_of_unity(cls, n: int):
        """omega_n = alpha^((order-1)/n) — FieldArray.primitive_root_of_unity, _fields/_array.py:1127-1187."""
        if not isinstance(n, (int, np.integer)):
            raise TypeError(f"Argument 'n' must be an int, not {type(n)}.")
        if not 1 <= n < cls.order:
            raise ValueError(f"Argument 'n' must be in [1, {cls.order}), not {n}.")
        if (cls.order - 1) % n != 0:
            raise ValueError(f"There are no primitive {n}-th roots of unity in {cls.name}.")
        e = (cls.order - 1) // n
        if cls._is_prime_field:
            return cls(pow(int(cls.primitive_element), e, cls.characteristic))
        return cls(_nt.powmod2(int(cls.primitive_element), e, int(cls.irreducible_poly)))

    @classmethod
    def compile(cls, mode: str):
        """The reference switches Numba targets here (_domains/_array.py:322-362); only 'cuda'/'auto' exist."""
        if mode not in ("auto", "cuda"):
            raise ValueError(f"Argument 'mode' must be in ['auto', 'cuda'] for the CUDA backend, not {mode!r}.")
        cls.ufunc_mode = "cuda"

    # ------------------------------------------------------------------------------ backend glue
    @classmethod
    def _handle(cls):
        """gfb_field* for the current CUDA device (created on first use)."""
        lib = _lib.lib()
        dev = 0
        try:
            import torch

            if torch.cuda.is_available():
                dev = torch.cuda.current_device()
        except ImportError:  # torch is optional for the host path
            pass
        with cls._lock:
            h = cls._handles.get(dev)
            if h is None:
                out = ctypes.c_void_p()
                poly = int(cls.irreducible_poly)
                if cls._is_prime_field:
                    st = lib.gfb_field_create(ctypes.byref(out), _lib.GFB_PRIME, _as_u64(cls.order), 1, 0, 0,
                                              _as_u64(int(cls.primitive_element)))
                else:
                    st = lib.gfb_field_create(ctypes.byref(out), _lib.GFB_BINARY, _as_u64(0), cls.degree,
                                              _as_u64(poly), _as_u64(poly >> 64),
                                              _as_u64(int(cls.primitive_element)))
                _lib.check(st, f"gfb_field_create({cls.name})")
                h = out.value
                cls._handles[dev] = h
        return h

    @classmethod
    def _elem_bytes(cls, dtype) -> int:
        return 16 if cls._wide else np.dtype(dtype).itemsize

    # -- wide fields: object ints <-> [n, 2] uint64 limbs (vectorised packer)
    @staticmethod
    def _pack_limbs(arr: np.ndarray) -> np.ndarray:
        flat = arr.ravel()
        raw = b"".join(int(v).to_bytes(16, "little") for v in flat)
        out = _pinned.empty((flat.size, 2), np.uint64)
        out[...] = np.frombuffer(raw, dtype=np.uint64).reshape(-1, 2)
        return out

    @staticmethod
    def _unpack_limbs(limbs: np.ndarray, shape) -> np.ndarray:
        raw = np.ascontiguousarray(limbs).tobytes()
        out = np.empty(limbs.shape[0], dtype=object)
        for i in range(limbs.shape[0]):
            out[i] = int.from_bytes(raw[16 * i: 16 * i + 16], "little")
        return out.reshape(shape)

    def to_limbs(self) -> np.ndarray:
        """[..., 2] uint64 (lo, hi) limbs of a wide-field array."""
        if not type(self)._wide:
            raise TypeError("to_limbs() is only defined for fields with 64 < degree <= 128.")
        return self._pack_limbs(self.view(np.ndarray)).reshape(self.shape + (2,))

    @classmethod
    def from_limbs(cls, limbs: np.ndarray):
        limbs = np.asarray(limbs, dtype=np.uint64)
        return cls._unpack_limbs(limbs.reshape(-1, 2), limbs.shape[:-1]).view(cls)

    # ------------------------------------------------------------------------------ ufunc dispatch
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        field = type(self)
        operands = list(range(len(inputs)))
        if method in ("at", "reduceat"):
            operands.pop(1)
        non_field = [i for i in operands if not isinstance(inputs[i], field)]
        dtype = self.dtype

        if ufunc in _OPS_BINARY or ufunc in (np.power, np.divmod, np.remainder) or ufunc in _OPS_UNARY or ufunc in (np.matmul, np.log, np.sqrt):
            return field._dispatch(ufunc, method, inputs, kwargs, non_field, operands, dtype)
        if ufunc in _UNSUPPORTED_UFUNCS:
            raise NotImplementedError(f"The NumPy ufunc {ufunc.__name__!r} is not supported on {field.name} arrays.")

        # everything else (comparisons, bitwise ops ...) runs on plain ndarray views, as in the reference
        v_inputs = [x.view(np.ndarray) if isinstance(x, FieldArray) else x for x in inputs]
        if "out" in kwargs:
            kwargs["out"] = tuple(o.view(np.ndarray) if isinstance(o, FieldArray) else o for o in kwargs["out"])
        output = getattr(ufunc, method)(*v_inputs, **kwargs)
        if ufunc in _UFUNCS_REQUIRING_VIEW and output is not None:
            output = field._view(output) if not np.isscalar(output) else field(output, dtype=dtype)
        return output

    def __pow__(self, other):
        return np.power(self, other)  # _domains/_ufunc.py:715-719

    @classmethod
    def _dispatch(cls, ufunc, method, inputs, kwargs, non_field, operands, dtype):
        out = kwargs.pop("out", None)
        kwargs.pop("casting", None)
        if kwargs.get("where", True) is not True:
            raise NotImplementedError("The 'where' keyword is not supported by the CUDA backend.")
        kwargs.pop("where", None)

        if ufunc in _OPS_UNARY:
            if method == "at":
                return cls._at("unary", _OPS_UNARY[ufunc], inputs[0], inputs[1], None)
            if method != "__call__":
                raise ValueError(f"Ufunc method {method!r} is not supported on {ufunc.__name__!r}. Reduction methods are only supported on binary functions.")
            result = cls._unary(_OPS_UNARY[ufunc], inputs[0], dtype)
        elif ufunc is np.power:
            if method in ("reduce", "accumulate", "reduceat"):
                raise ValueError(f"Ufunc method {method!r} is not supported on 'power' because it takes inputs with type {cls.name} array and integer array.")
            if method == "at":
                return cls._at("power", None, inputs[0], inputs[1], inputs[2])
            result = cls._power(inputs, operands, dtype, outer=(method == "outer"))
        elif ufunc in (np.divmod, np.remainder):
            if non_field:
                raise TypeError(f"Operation {ufunc.__name__!r} requires both operands to be instances of {cls.name}.")
            if ufunc is np.divmod:
                q = np.true_divide(*inputs)
                return q, cls.Zeros(q.shape, dtype=dtype)
            x = np.add(*inputs)
            return cls.Zeros(x.shape, dtype=dtype)
        elif ufunc is np.matmul:
            if method != "__call__":
                raise ValueError(f"Ufunc method {method!r} is not supported on 'matmul'. Only '__call__' is supported.")
            from ._linalg import matmul

            result = matmul(cls, inputs[0], inputs[1])
        elif ufunc in (np.log, np.sqrt):
            raise NotImplementedError(f"{ufunc.__name__!r} is outside the CUDA backend's hot path (SURVEY.md §8).")
        else:
            op = _OPS_BINARY[ufunc]
            if ufunc is np.multiply and non_field:
                # scalar multiplication by integers: reduce them modulo the characteristic (_ufunc.py:392-401)
                inputs = list(inputs)
                for i in non_field:
                    v = inputs[i]
                    if isinstance(v, (int, np.integer)):
                        inputs[i] = cls(int(v) % cls.characteristic, dtype=dtype)
                    elif isinstance(v, np.ndarray) and (np.issubdtype(v.dtype, np.integer) or v.dtype == np.object_):
                        inputs[i] = cls(np.mod(v, cls.characteristic), dtype=dtype)
                    else:
                        raise TypeError(f"Operation 'multiply' requires operands that are not {cls.name} arrays to be integers or an integer np.ndarray, not {type(v)}.")
            elif non_field:
                raise TypeError(
                    f"Operation {ufunc.__name__!r} requires both operands to be instances of {cls.name}, "
                    f"not {[type(inputs[i]) for i in operands]}."
                )
            if method == "__call__":
                result = cls._binary(op, inputs[0], inputs[1], dtype)
            elif method == "outer":
                a, b = np.asarray(inputs[0]), np.asarray(inputs[1])
                a2 = a.reshape(a.shape + (1,) * b.ndim)
                result = cls._binary(op, a2.view(cls), b.view(cls), dtype)
            elif method == "reduce":
                result = cls._reduce(op, inputs[0], kwargs.get("axis", 0), kwargs.get("keepdims", False), dtype)
            elif method == "accumulate":
                result = cls._accumulate(op, inputs[0], kwargs.get("axis", 0), dtype)
            elif method == "reduceat":
                result = cls._reduceat(op, inputs[0], inputs[1], kwargs.get("axis", 0), dtype)
            elif method == "at":
                return cls._at("binary", op, inputs[0], inputs[1], inputs[2] if len(inputs) > 2 else None)
            else:
                raise ValueError(f"Unknown ufunc method {method!r}.")

        if out is not None:
            target = out[0] if isinstance(out, tuple) else out
            target[...] = result
            return target
        return result

    # host kernels ----------------------------------------------------------------------------------
    @classmethod
    def _prep(cls, x, dtype):
        """ndarray view of a field operand in the working dtype, C-contiguous."""
        a = x.view(np.ndarray) if isinstance(x, np.ndarray) else np.asarray(x)
        if cls._wide:
            return a
        if a.dtype != dtype:
            a = a.astype(dtype)
        return a

    @classmethod
    def _to_dev_layout(cls, a: np.ndarray) -> np.ndarray:
        return cls._pack_limbs(a) if cls._wide else np.ascontiguousarray(a)

    @classmethod
    def _from_dev_layout(cls, out: np.ndarray, shape, dtype):
        if cls._wide:
            return cls._unpack_limbs(out, shape).view(cls)
        return out.reshape(shape).view(cls)

    @classmethod
    def _alloc_out(cls, n: int, dtype):
        return _pinned.empty((n, 2), np.uint64) if cls._wide else _pinned.empty(n, dtype)

    @classmethod
    def _scalar_limbs(cls, v: int):
        return _as_u64(v), _as_u64(v >> 64)

    @classmethod
    def _binary(cls, op, a, b, dtype):
        lib = _lib.lib()
        h = cls._handle()
        a = cls._prep(a, dtype)
        b = cls._prep(b, dtype)
        shape = np.broadcast_shapes(a.shape, b.shape)
        n = int(np.prod(shape)) if shape else 1
        eb = cls._elem_bytes(dtype)
        out = cls._alloc_out(n, dtype)
        if n == 0:
            return cls._from_dev_layout(out, shape, dtype)
        if b.size == 1 and a.size == n:
            s_lo, s_hi = cls._scalar_limbs(int(b.reshape(-1)[0]))
            src = cls._to_dev_layout(a)
            st = lib.gfb_binary_scalar_host(h, op, eb, src.ctypes.data, s_lo, s_hi, 0, out.ctypes.data, n)
        elif a.size == 1 and b.size == n:
            s_lo, s_hi = cls._scalar_limbs(int(a.reshape(-1)[0]))
            src = cls._to_dev_layout(b)
            st = lib.gfb_binary_scalar_host(h, op, eb, src.ctypes.data, s_lo, s_hi, 1, out.ctypes.data, n)
        else:
            aa = cls._to_dev_layout(np.broadcast_to(a, shape))
            bb = cls._to_dev_layout(np.broadcast_to(b, shape))
            st = lib.gfb_binary_host(h, op, eb, aa.ctypes.data, bb.ctypes.data, out.ctypes.data, n)
        _lib.check(st, "gfb_binary_host")
        return cls._from_dev_layout(out, shape, dtype)

    @classmethod
    def _unary(cls, op, a, dtype):
        lib = _lib.lib()
        h = cls._handle()
        a = cls._prep(a, dtype)
        shape = a.shape
        n = a.size
        out = cls._alloc_out(n, dtype)
        if n:
            src = cls._to_dev_layout(a)
            st = lib.gfb_unary_host(h, op, cls._elem_bytes(dtype), src.ctypes.data, out.ctypes.data, n)
            _lib.check(st, "gfb_unary_host")
        return cls._from_dev_layout(out, shape, dtype)

    @classmethod
    def _power(cls, inputs, operands, dtype, outer=False):
        base, expo = inputs[0], inputs[1]
        if not isinstance(base, cls):
            raise TypeError(f"Operation 'power' requires the first operand to be a {cls.name} array, not {type(base)}.")
        if isinstance(expo, FieldArray):
            raise TypeError(f"Operation 'power' requires the second operand to be an integer array, not {type(expo)}.")
        if isinstance(expo, (int, np.integer)):
            pass
        elif isinstance(expo, np.ndarray):
            if not (np.issubdtype(expo.dtype, np.integer) or (cls._wide and expo.dtype == np.object_)):
                raise ValueError(f"Operation 'power' requires operands with type np.ndarray to have integer dtype, not {expo.dtype}.")
        else:
            raise TypeError(f"Operation 'power' requires the second operand to be an integer or integer np.ndarray, not {type(expo)}.")
        lib = _lib.lib()
        h = cls._handle()
        a = cls._prep(base, dtype)
        eb = cls._elem_bytes(dtype)
        if isinstance(expo, (int, np.integer)) or np.asarray(expo).ndim == 0:
            e = int(expo)
            if not -(2**63) <= e < 2**63:
                e = cls._reduce_exponent(e)
            n = a.size
            out = cls._alloc_out(n, dtype)
            if n:
                src = cls._to_dev_layout(a)
                st = lib.gfb_power_scalar_host(h, eb, src.ctypes.data, ctypes.c_int64(e), out.ctypes.data, n)
                _lib.check(st, "gfb_power_scalar_host")
            return cls._from_dev_layout(out, a.shape, dtype)
        e = np.asarray(expo)
        if outer:
            a = a.reshape(a.shape + (1,) * e.ndim)
        if e.dtype == np.object_:
            e = np.array([cls._reduce_exponent(int(v)) for v in e.ravel()], dtype=np.int64).reshape(e.shape)
        shape = np.broadcast_shapes(a.shape, e.shape)
        n = int(np.prod(shape)) if shape else 1
        out = cls._alloc_out(n, dtype)
        if n:
            aa = cls._to_dev_layout(np.broadcast_to(a, shape))
            ee = np.ascontiguousarray(np.broadcast_to(e, shape), dtype=np.int64)
            st = lib.gfb_power_host(h, eb, aa.ctypes.data, ee.ctypes.data, out.ctypes.data, n)
            _lib.check(st, "gfb_power_host")
        return cls._from_dev_layout(out, shape, dtype)

    @classmethod
    def _reduce_exponent(cls, e: int) -> int:
        """Bring an arbitrary Python-int exponent into int64 range without changing a**e (a != 0)."""
        if -(2**63) <= e < 2**63:
            return e
        q1 = cls.order - 1
        if e < 0:
            # keep the sign: 0 ** negative must still raise ZeroDivisionError in the kernel
            r = (-e) % q1
            return -(r if r else q1)
        r = e % q1
        return r if r else q1

    # -- ufunc methods (reduce / accumulate / reduceat / at): _domains/_ufunc.py:349,403 forwards them to the DUFunc;
    #    here they are gfb_fold / gfb_fold_segments (csrc/gfb_fold.cu) plus the element-wise kernels
    @classmethod
    def _fold(cls, inner_op, x, axis, scan, dtype):
        """Fold (scan=False) or running fold (scan=True) of ndarray ``x`` along ``axis`` with ADD or MUL."""
        lib = _lib.lib()
        x = np.ascontiguousarray(cls._prep(x, dtype))
        axis = axis % x.ndim
        moved = False
        outer = int(np.prod(x.shape[:axis])) if axis else 1
        length = x.shape[axis]
        inner = int(np.prod(x.shape[axis + 1:])) if axis + 1 < x.ndim else 1
        if inner != 1 and outer * inner < 2048 and length >= 8192:
            # few long strided lines: make them contiguous so the chunked kernels apply
            x = np.ascontiguousarray(np.moveaxis(x, axis, -1))
            moved, outer, inner = True, outer * inner, 1
        out_shape = x.shape if scan else (x.shape[:-1] if moved else x.shape[:axis] + x.shape[axis + 1:])
        out = _pinned.empty(out_shape, dtype)
        if out.size:
            st = lib.gfb_fold_host(cls._handle(), inner_op, 1 if scan else 0, cls._elem_bytes(dtype), x.ctypes.data,
                                   out.ctypes.data, outer, length, inner)
            _lib.check(st, "gfb_fold_host")
        if moved and scan:
            out = np.moveaxis(out, -1, axis)
        return out

    @classmethod
    def _reduce(cls, op, a, axis, keepdims, dtype):
        """np.add.reduce / np.multiply.reduce, and the left folds of subtract / divide (a0 - a1 - ... = a0 - sum(rest))."""
        x = np.asarray(a.view(np.ndarray))
        if axis is None:
            x = x.reshape(-1)
            axis = 0
        if isinstance(axis, tuple):
            res = x.view(cls)
            for ax in sorted((v % x.ndim for v in axis), reverse=True):
                res = cls._reduce(op, np.atleast_1d(res).view(cls), ax, keepdims, dtype)
            return res
        axis = axis % max(x.ndim, 1)
        if x.shape[axis] == 0:
            raise ValueError("zero-size array to reduction operation which has no identity")
        inner_op = _lib.GFB_ADD if op in (_lib.GFB_ADD, _lib.GFB_SUB) else _lib.GFB_MUL
        if cls._wide:
            xm = np.moveaxis(x, axis, 0)
            if op in (_lib.GFB_SUB, _lib.GFB_DIV) and xm.shape[0] > 1:
                res = cls._binary(op, xm[0].view(cls), cls._tree(inner_op, xm[1:], dtype), dtype)
            else:
                res = cls._tree(inner_op, xm, dtype)
            res = np.asarray(res)
        elif op in (_lib.GFB_SUB, _lib.GFB_DIV) and x.shape[axis] > 1:
            head = np.take(x, 0, axis=axis)
            rest = cls._fold(inner_op, x[(slice(None),) * axis + (slice(1, None),)], axis, False, dtype)
            res = np.asarray(cls._binary(op, np.asarray(head).view(cls), np.asarray(rest).view(cls), dtype))
        else:
            res = cls._fold(inner_op, x, axis, False, dtype)
        if keepdims:
            res = np.expand_dims(res, axis)
        res = np.asarray(res)
        if res.ndim == 0:
            return cls(int(res.item()), dtype=dtype)
        return res.view(cls)

    @classmethod
    def _tree(cls, op, x, dtype):
        cur = x
        while cur.shape[0] > 1:
            half = cur.shape[0] // 2
            head = cls._binary(op, np.ascontiguousarray(cur[:half]).view(cls), np.ascontiguousarray(cur[half: 2 * half]).view(cls), dtype)
            head = head.view(np.ndarray)
            if cur.shape[0] % 2:
                head = np.concatenate([head, cur[2 * half:]], axis=0)
            cur = head
        return cur[0].view(cls) if isinstance(cur[0], np.ndarray) else cls(cur[0], dtype=dtype)

    @classmethod
    def _accumulate(cls, op, a, axis, dtype):
        """np.<ufunc>.accumulate: out[k] = a0 (op) a1 (op) ... (op) ak along ``axis``."""
        if cls._wide:
            raise NotImplementedError(f"ufunc.accumulate is not implemented over {cls.name} (elements wider than 64 bits).")
        x = np.asarray(a.view(np.ndarray))
        if x.ndim == 0:
            raise ValueError("cannot accumulate on a scalar")
        axis = axis % x.ndim
        inner_op = _lib.GFB_ADD if op in (_lib.GFB_ADD, _lib.GFB_SUB) else _lib.GFB_MUL
        if x.shape[axis] == 0:
            return x.astype(dtype).view(cls)
        if op in (_lib.GFB_ADD, _lib.GFB_MUL) or x.shape[axis] == 1:
            return np.asarray(cls._fold(inner_op, x, axis, True, dtype)).view(cls)
        # a0 - (a1 + ... + ak): running fold of the tail, then one element-wise pass against the broadcast head
        tail = x[(slice(None),) * axis + (slice(1, None),)]
        run = np.asarray(cls._fold(inner_op, tail, axis, True, dtype))
        head = np.take(x, [0], axis=axis)
        rest = np.asarray(cls._binary(op, np.broadcast_to(head, run.shape).view(cls), run.view(cls), dtype))
        return np.concatenate([head.astype(dtype), rest], axis=axis).view(cls)

    @classmethod
    def _reduceat(cls, op, a, indices, axis, dtype):
        """np.<ufunc>.reduceat (NumPy semantics: segment i is a[idx[i]:idx[i+1]], or the single element a[idx[i]] when
        idx[i] >= idx[i+1]; the last segment runs to the end)."""
        if cls._wide:
            raise NotImplementedError(f"ufunc.reduceat is not implemented over {cls.name} (elements wider than 64 bits).")
        x = np.asarray(a.view(np.ndarray))
        idx = np.asarray(indices)
        if idx.ndim != 1 or not np.issubdtype(idx.dtype, np.integer):
            raise TypeError("Argument 'indices' of reduceat must be a 1-D integer array.")
        axis = axis % x.ndim
        n = x.shape[axis]
        idx = idx.astype(np.int64)
        if idx.size and (idx.min() < 0 or idx.max() >= n):
            raise IndexError(f"index {int(idx.max() if idx.max() >= n else idx.min())} out-of-bounds in {('add', 'subtract', 'multiply', 'divide')[op]}.reduceat [0, {n})")
        xm = np.ascontiguousarray(np.moveaxis(cls._prep(x, dtype), axis, -1))  # [lines][n]
        lines = xm.size // n if n else 0
        nseg = idx.size
        end = np.concatenate([idx[1:], [n]]) if nseg else idx
        end = np.where(end > idx, end, idx + 1)  # single-element segments
        inner_op = _lib.GFB_ADD if op in (_lib.GFB_ADD, _lib.GFB_SUB) else _lib.GFB_MUL
        out = _pinned.empty(xm.shape[:-1] + (nseg,), dtype)
        if out.size:
            offs = (np.arange(lines, dtype=np.int64) * n)[:, None]
            lib = _lib.lib()
            if op in (_lib.GFB_ADD, _lib.GFB_MUL):
                start_all = np.ascontiguousarray((offs + idx[None, :]).reshape(-1))
                end_all = np.ascontiguousarray((offs + end[None, :]).reshape(-1))
                st = lib.gfb_fold_segments_host(cls._handle(), inner_op, cls._elem_bytes(dtype), xm.ctypes.data, xm.size,
                                                start_all.ctypes.data, end_all.ctypes.data, out.ctypes.data, start_all.size)
                _lib.check(st, "gfb_fold_segments_host")
            else:
                # left fold a[s] (op) (a[s+1] + ... ): fold the tails of the segments that have one, then one pass
                flat = xm.reshape(lines, n)
                heads = flat[:, idx]
                res = heads.copy()
                has_tail = end - idx > 1
                if has_tail.any():
                    ts = np.ascontiguousarray((offs + idx[None, has_tail] + 1).reshape(-1))
                    te = np.ascontiguousarray((offs + end[None, has_tail]).reshape(-1))
                    tails = np.empty(ts.size, dtype=dtype)
                    st = lib.gfb_fold_segments_host(cls._handle(), inner_op, cls._elem_bytes(dtype), xm.ctypes.data, xm.size,
                                                    ts.ctypes.data, te.ctypes.data, tails.ctypes.data, ts.size)
                    _lib.check(st, "gfb_fold_segments_host")
                    hh = np.ascontiguousarray(heads[:, has_tail]).reshape(-1)
                    res[:, has_tail] = np.asarray(cls._binary(op, hh.view(cls), tails.view(cls), dtype)).reshape(lines, -1)
                out[...] = res.reshape(out.shape)
        return np.moveaxis(np.asarray(out), -1, axis).view(cls)

    @classmethod
    def _at(cls, kind, op, a, indices, b):
        """np.<ufunc>.at(a, indices[, b]): unbuffered in-place a[indices] = op(a[indices], b).  An index that occurs k
        times is updated k times, in order: the selected positions are processed in rounds (round r takes the r-th
        occurrence of every index, so positions are unique inside a round) of gather -> element-wise kernel -> scatter."""
        if not isinstance(a, cls):
            raise TypeError(f"The first operand of ufunc.at must be a {cls.name} array, not {type(a)}.")
        dtype = a.dtype
        flat_index = np.arange(a.size, dtype=np.int64).reshape(a.shape)[indices]
        pos = np.atleast_1d(np.asarray(flat_index)).reshape(-1)
        if kind != "unary":
            if b is None:
                raise ValueError("second operand needed for ufunc")
            if kind == "binary":
                if isinstance(b, (int, np.integer)) and op == _lib.GFB_MUL:
                    b = cls(int(b) % cls.characteristic, dtype=dtype)
                elif not isinstance(b, cls):
                    raise TypeError(f"Operation requires both operands to be instances of {cls.name}, not {type(b)}.")
            elif isinstance(b, FieldArray):
                raise TypeError(f"Operation 'power' requires the second operand to be an integer array, not {type(b)}.")
            bb = np.broadcast_to(np.asarray(b.view(np.ndarray) if isinstance(b, np.ndarray) else b), np.shape(flat_index)).reshape(-1)
        if pos.size == 0:
            return None
        order = np.argsort(pos, kind="stable")
        sp = pos[order]
        first = np.concatenate([[True], sp[1:] != sp[:-1]])
        group_start = np.maximum.accumulate(np.where(first, np.arange(sp.size), 0))
        rank = np.arange(sp.size) - group_start
        target = a.view(np.ndarray)
        if not target.flags.c_contiguous:
            raise ValueError("ufunc.at on the CUDA backend needs a C-contiguous array.")
        flat = target.reshape(-1)
        for r in range(int(rank.max()) + 1):
            sel = order[rank == r]
            p = pos[sel]
            vals = np.ascontiguousarray(flat[p]).view(cls)
            if kind == "unary":
                res = cls._unary(op, vals, dtype)
            elif kind == "power":
                res = cls._power([vals, np.ascontiguousarray(bb[sel])], [0, 1], dtype)
            else:
                res = cls._binary(op, vals, np.ascontiguousarray(bb[sel]).view(cls) if not cls._wide else bb[sel].view(cls), dtype)
            flat[p] = np.asarray(res).reshape(-1)
        return None

    # ------------------------------------------------------------------------------ array functions
    def __array_function__(self, func, types, args, kwargs):
        field = type(self)
        if func is np.convolve:
            from ._linalg import convolve

            return convolve(field, *args, **kwargs)
        if func is np.fft.fft or func is np.fft.ifft:
            from ._ntt import _fft_dispatch

            return _fft_dispatch(field, func is np.fft.ifft, *args, **kwargs)
        name = _LINALG_FUNCTIONS.get(func)
        if name is not None:
            from . import _linalg

            return getattr(_linalg, name)(field, *args, **kwargs)
        return super().__array_function__(func, types, args, kwargs)

    # ------------------------------------------------------------------------------ linear algebra methods
    def row_reduce(self, ncols=None, eye="left"):
        """Reduced row echelon form (FieldArray.row_reduce, _fields/_array.py:1412-1469)."""
        from . import _linalg

        return _linalg.row_reduce(self, ncols=ncols, eye=eye)

    def lu_decompose(self):
        from . import _linalg

        return _linalg.lu_decompose(self)

    def plu_decompose(self):
        from . import _linalg

        return _linalg.plu_decompose(self)

    def row_space(self):
        from . import _linalg

        return _linalg.row_space(self)

    def column_space(self):
        from . import _linalg

        return _linalg.column_space(self)

    def left_null_space(self):
        from . import _linalg

        return _linalg.left_null_space(self)

    def null_space(self):
        from . import _linalg

        return _linalg.null_space(self)

    # ------------------------------------------------------------------------------ misc
    def to_device(self):
        return DeviceArray.from_host(self)

    @classmethod
    def device(cls, x, dtype=None):
        return DeviceArray.from_host(x if isinstance(x, cls) else cls(x, dtype=dtype))

    def __repr__(self):
        body = np.array2string(self.view(np.ndarray), separator=", ", prefix="GF(")
        return f"GF({body}, order={type(self).order_str})"

    def __str__(self):
        return np.array2string(self.view(np.ndarray))


# -------------------------------------------------------------------------------------------------
# Device-resident arrays
# -------------------------------------------------------------------------------------------------
class DeviceArray:
    """
    A field array whose storage is a ``torch`` CUDA tensor (torch is only the allocator/stream
    provider).  Supports the same arithmetic operators as ``FieldArray``; operands must have equal
    shapes or be scalars (no general broadcasting on the device).  ``.cpu()`` → ``FieldArray``.
    """

    __array_priority__ = 1000

    def __init__(self, field, tensor, shape, dtype):
        self.field = field
        self.tensor = tensor  # flat uint8 storage
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    @property
    def size(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def elem_bytes(self) -> int:
        return self.field._elem_bytes(self.dtype)

    @property
    def data_ptr(self) -> int:
        return self.tensor.data_ptr()

    @classmethod
    def from_host(cls, x: FieldArray):
        import torch

        field = type(x)
        _lib.lib()
        host = field._to_dev_layout(x.view(np.ndarray))
        t = torch.from_numpy(host.view(np.uint8).reshape(-1)).to("cuda", non_blocking=False)
        return cls(field, t, x.shape, x.dtype)

    @classmethod
    def empty(cls, field, shape, dtype=None):
        import torch

        dtype = field._get_dtype(dtype)
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = int(np.prod(shape)) if shape else 1
        t = torch.empty(n * field._elem_bytes(dtype), dtype=torch.uint8, device="cuda")
        return cls(field, t, shape, dtype)

    def cpu(self) -> FieldArray:
        host = self.tensor.cpu().numpy()
        if self.field._wide:
            return self.field._unpack_limbs(host.view(np.uint64).reshape(-1, 2), self.shape).view(self.field)
        return host.view(self.dtype).reshape(self.shape).view(self.field)

    @staticmethod
    def _stream():
        import torch

        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check_errors(self):
        """Raise ZeroDivisionError if a kernel flagged one (synchronises the current stream)."""
        st = _lib.lib().gfb_take_error(self.field._handle(), self._stream())
        _lib.check(st, "gfb_take_error")

    def _binary(self, other, op, reverse=False, check=None):
        lib = _lib.lib()
        f = self.field
        h = f._handle()
        out = DeviceArray.empty(f, self.shape, self.dtype)
        n = self.size
        if isinstance(other, DeviceArray):
            if other.field is not f:
                raise TypeError(f"Both operands must be arrays over {f.name}.")
            if other.shape != self.shape or other.dtype != self.dtype:
                raise ValueError("DeviceArray operands must have identical shape and dtype (or be scalars).")
            a, b = (other, self) if reverse else (self, other)
            st = lib.gfb_binary(h, op, self.elem_bytes, a.data_ptr, b.data_ptr, out.data_ptr, n, self._stream())
        else:
            if isinstance(other, FieldArray):
                if type(other) is not f or other.size != 1:
                    raise TypeError("Scalar operands must be 0-d arrays over the same field.")
                v = int(other.reshape(-1)[0])
            elif isinstance(other, (int, np.integer)) and op == _lib.GFB_MUL:
                v = int(other) % f.characteristic
            else:
                raise TypeError(f"Operation requires both operands to be arrays over {f.name}, not {type(other)}.")
            lo, hi = f._scalar_limbs(v)
            st = lib.gfb_binary_scalar(h, op, self.elem_bytes, self.data_ptr, lo, hi, 1 if reverse else 0,
                                       out.data_ptr, n, self._stream())
        _lib.check(st, "gfb_binary")
        if op == _lib.GFB_DIV and (check is None or check):
            out._check_errors()
        return out

    def _unary(self, op):
        lib = _lib.lib()
        out = DeviceArray.empty(self.field, self.shape, self.dtype)
        st = lib.gfb_unary(self.field._handle(), op, self.elem_bytes, self.data_ptr, out.data_ptr, self.size, self._stream())
        _lib.check(st, "gfb_unary")
        if op == _lib.GFB_INV:
            out._check_errors()
        return out

    def __add__(self, o):
        return self._binary(o, _lib.GFB_ADD)

    def __radd__(self, o):
        return self._binary(o, _lib.GFB_ADD, reverse=True)

    def __sub__(self, o):
        return self._binary(o, _lib.GFB_SUB)

    def __rsub__(self, o):
        return self._binary(o, _lib.GFB_SUB, reverse=True)

    def __mul__(self, o):
        return self._binary(o, _lib.GFB_MUL)

    def __rmul__(self, o):
        return self._binary(o, _lib.GFB_MUL, reverse=True)

    def __truediv__(self, o):
        return self._binary(o, _lib.GFB_DIV)

    def __rtruediv__(self, o):
        return self._binary(o, _lib.GFB_DIV, reverse=True)

    __floordiv__ = __truediv__

    def __neg__(self):
        return self._unary(_lib.GFB_NEG)

    def reciprocal(self):
        return self._unary(_lib.GFB_INV)

    def square(self):
        return self._unary(_lib.GFB_SQUARE)

    def __pow__(self, e):
        lib = _lib.lib()
        out = DeviceArray.empty(self.field, self.shape, self.dtype)
        h = self.field._handle()
        if isinstance(e, (int, np.integer)):
            e = self.field._reduce_exponent(int(e))
            st = lib.gfb_power_scalar(h, self.elem_bytes, self.data_ptr, ctypes.c_int64(e), out.data_ptr, self.size, self._stream())
        else:
            import torch

            if isinstance(e, np.ndarray):
                e = torch.from_numpy(np.ascontiguousarray(e, dtype=np.int64)).cuda()
            if not (isinstance(e, torch.Tensor) and e.dtype == torch.int64 and e.numel() == self.size):
                raise TypeError("Exponent must be an int or an int64 array/tensor with the same number of elements.")
            st = lib.gfb_power(h, self.elem_bytes, self.data_ptr, e.data_ptr(), out.data_ptr, self.size, self._stream())
        _lib.check(st, "gfb_power")
        out._check_errors()
        return out

    def __repr__(self):
        return f"DeviceArray({self.field.name}, shape={self.shape}, dtype={self.dtype.name}, device=cuda)"


# -------------------------------------------------------------------------------------------------
# Factory
# -------------------------------------------------------------------------------------------------
_CLASS_CACHE: dict = {}
_CACHE_LOCK = threading.Lock()


def _valid_dtypes(order: int) -> list:
    # _fields/_ufunc.py:36-48 (prime fields need (p-1)^2 to fit) and _domains/_meta.py
    return [d for d in _DTYPES if np.iinfo(d).max >= order - 1]


def GF(order, *, irreducible_poly=None, primitive_element=None, verify=True, compile=None, repr=None):  # noqa: A002
    """
    Field-class factory mirroring ``galois.GF`` (_fields/_factory.py:52-301) for the field families on the hot
    path: prime fields GF(p) with p < 2^32 and binary extension fields GF(2^m) with m <= 128.

    ``GF(p**m)`` or ``GF(p, m)`` styles are both accepted; classes are flyweights keyed by
    (order, irreducible_poly, primitive_element) like the reference's (_factory.py:385,486).
    """
    if isinstance(order, (np.integer,)):
        order = int(order)
    if not isinstance(order, int):
        raise TypeError(f"Argument 'order' must be an integer, not {type(order)}.")
    if compile not in (None, "auto", "cuda"):
        raise ValueError(f"Argument 'compile' must be in ['auto', 'cuda'] for the CUDA backend, not {compile!r}.")
    p, m = _nt.prime_power(order)

    if m == 1:
        if irreducible_poly is not None and verify:
            pass  # degree-1 polynomial x - alpha; the reference accepts it, nothing to check here
        if primitive_element is None:
            alpha = _nt.primitive_root(p)
        else:
            alpha = int(primitive_element)
            if verify and not _nt.is_primitive_root(alpha, p):
                raise ValueError(f"Argument 'primitive_element' must be a primitive root modulo {p}, {alpha} is not.")
        if p >= 2**32:
            raise NotImplementedError(
                f"GF({p}): primes >= 2^32 run in the reference's object-dtype 'python-calculate' mode and are "
                "outside the CUDA hot path (SURVEY.md §8a)."
            )
        poly = p + ((p - alpha) % p)  # x - alpha, integer representation (the reference's GF(p).irreducible_poly)
        key = (p, 1, alpha)
    else:
        if p != 2:
            raise NotImplementedError(
                f"GF({p}^{m}): odd-characteristic extension fields are outside the CUDA hot path (SURVEY.md §2.2)."
            )
        if m > 128:
            raise NotImplementedError(f"GF(2^{m}): degrees above 128 are not supported by the CUDA backend.")
        if irreducible_poly is None:
            if m not in _nt.CONWAY_2:
                raise LookupError(
                    f"The Conway polynomial C_{{2,{m}}} is not available. Specify `irreducible_poly` explicitly."
                )
            poly = _nt.CONWAY_2[m]
            alpha = 2 if primitive_element is None else int(primitive_element)
        else:
            poly = _nt.parse_poly2(irreducible_poly)
            if poly.bit_length() - 1 != m:
                raise ValueError(f"Argument 'irreducible_poly' must have degree equal to {m}, not {poly.bit_length() - 1}.")
            if verify and not _nt.is_irreducible2(poly):
                raise ValueError(f"Argument 'irreducible_poly' must be irreducible, {poly:#x} is not.")
            if primitive_element is None:
                alpha = 2 if not verify and m > 20 else _nt.smallest_primitive_element2(poly)
            else:
                alpha = int(primitive_element)
        if primitive_element is not None and verify and m <= 64 and not _nt.is_primitive_element2(alpha, poly):
            raise ValueError(f"Argument 'primitive_element' must be a multiplicative generator of GF(2^{m}), {alpha} is not.")
        key = (2, m, alpha, poly)

    with _CACHE_LOCK:
        cls = _CLASS_CACHE.get(key)
        if cls is not None:
            return cls
        wide = m >= 64  # order - 1 no longer fits int64: the reference switches to dtype=object (_domains/_meta.py)
        if wide:
            dtypes = [np.object_]
        else:
            # prime fields: the reference additionally requires (p-1)^2 <= int64 max for its int64 kernels and falls
            # back to dtype=object beyond (_fields/_ufunc.py:36-48); the CUDA kernels reduce in registers, so
            # primes up to 2^32 keep native dtypes here.
            dtypes = _valid_dtypes(order)
        name = f"GF({p})" if m == 1 else f"GF({p}^{m})"
        attrs = dict(
            name=name,
            characteristic=p,
            degree=m,
            order=order,
            order_str=str(p) if m == 1 else f"{p}^{m}",
            irreducible_poly=poly if poly is not None else None,
            primitive_element=alpha,
            dtypes=dtypes,
            _is_prime_field=(m == 1),
            _wide=wide,
            _handles={},
            _lock=threading.Lock(),
        )
        cls = FieldMeta(f"FieldArray_{p}_{m}_{alpha}", (FieldArray,), attrs)
        cls.__module__ = "galois_b200"
        cls.__qualname__ = name
        _CLASS_CACHE[key] = cls
        return cls


Field = GF
