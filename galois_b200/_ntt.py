"""
Number-theoretic transforms: ``ntt`` / ``intt`` and the ``np.fft.fft`` / ``np.fft.ifft`` overrides.

Mirrors /root/reference/src/galois/_ntt.py:16-278 (argument checking, modulus selection) and
``fft_jit.__call__`` (_domains/_function.py:177-212: zero-padding/truncation to ``n``, omega from
``primitive_root_of_unity``, omega^-1 for the inverse, division by N when ``_direction == norm``).
The transform itself is ``gfb_ntt_exec`` (include/galois_b200.h).  One addition over the reference,
which is 1-D only (_function.py:180-181): a 2-D input is a batch of independent row transforms.
"""

from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib, _nt, _pinned
from ._field import GF, DeviceArray, FieldArray

_PLANS: dict = {}
_PLAN_LOCK = threading.Lock()


def _plan(field, n: int, inverse: bool, scaled: bool):
    """gfb_ntt_plan* for (field, n, direction, scaling) on the current device."""
    lib = _lib.lib()
    h = field._handle()
    key = (field, h, n, inverse, scaled)
    with _PLAN_LOCK:
        plan = _PLANS.get(key)
        if plan is None:
            p = field.characteristic
            omega = int(field.primitive_root_of_unity(n)) if n > 1 else 1
            if inverse:  # omega ** -1 (_function.py:194-195); omega has order n, so omega^-1 = omega^(n-1)
                if field._is_prime_field:
                    omega = pow(omega, -1, p)
                else:
                    omega = _nt.powmod2(omega, n - 1, int(field.irreducible_poly))
            # y /= field(n % p) (_function.py:209-210); in characteristic 2 every admissible n is odd: n % 2 = 1
            scale = (pow(n % p, -1, p) if p > 2 else 1) if scaled else 1
            out = ctypes.c_void_p()
            st = lib.gfb_ntt_plan_create(ctypes.byref(out), h, n, ctypes.c_uint64(omega), ctypes.c_uint64(scale))
            _lib.check(st, f"gfb_ntt_plan_create(n={n})")
            plan = out.value
            _PLANS[key] = plan
    return plan


def _check_size(field, n: int):
    if n < 1:
        raise ValueError(f"Argument 'n' must be at least 1, not {n}.")
    if field._wide:
        raise NotImplementedError(f"The CUDA FFT is not implemented over {field.name} (degree above 63).")
    if (field.order - 1) % n != 0:
        raise ValueError(f"There are no primitive {n}-th roots of unity in {field.name}.")


def _fft_dispatch(field, inverse: bool, x, n=None, axis=-1, norm=None):
    """np.fft.fft / np.fft.ifft on a FieldArray — fft_jit.__call__, _domains/_function.py:177-212."""
    if not isinstance(x, field):
        raise TypeError(f"Argument 'x' must be a {field.name} array, not {type(x)}.")
    norm = "backward" if norm is None else norm
    if axis != -1:
        raise ValueError("The FFT is only implemented on 1-D arrays.")
    if norm not in ("forward", "backward"):
        raise ValueError("DFT normalization can only be applied to the forward or backward transform, not 'ortho'.")
    if x.ndim not in (1, 2):
        raise ValueError("The CUDA FFT accepts 1-D arrays or a 2-D batch of rows.")
    dtype = x.dtype
    length = x.shape[-1]
    if n is None:
        n = length
    elif n < length:
        x = x[..., :n]
    elif n > length:
        pad = field.Zeros(x.shape[:-1] + (n - length,), dtype=dtype)
        x = np.concatenate([x.view(np.ndarray), pad.view(np.ndarray)], axis=-1).view(field)
    _check_size(field, n)
    direction = "backward" if inverse else "forward"
    scaled = direction == norm
    plan = _plan(field, n, inverse, scaled)

    src = np.ascontiguousarray(x.view(np.ndarray))
    eb = src.dtype.itemsize
    if eb < 4:
        # the transform kernels carry 4- or 8-byte elements; small dtypes are widened on the host
        src = src.astype(np.uint32)
        eb = 4
    batch = 1 if src.ndim == 1 else src.shape[0]
    out = _pinned.empty(src.shape, src.dtype)
    st = _lib.lib().gfb_ntt_exec_host(plan, eb, src.ctypes.data, out.ctypes.data, batch)
    _lib.check(st, "gfb_ntt_exec_host")
    if out.dtype != dtype:
        out = out.astype(dtype)
    return out.view(field)


def _ntt(x, size=None, modulus=None, forward=True, scaled=True):
    """_ntt.py:239-278."""
    for name, v in (("size", size), ("modulus", modulus)):
        if v is not None and not isinstance(v, (int, np.integer)):
            raise TypeError(f"Argument {name!r} must be an int, not {type(v)}.")
    if size is None:
        size = np.shape(x)[-1] if isinstance(x, np.ndarray) else len(x)
    if isinstance(x, FieldArray) and modulus is not None and type(x).characteristic == modulus:
        xmax = 0  # elements of a GF(modulus) array are < modulus by construction: skip the O(n) scan
    else:
        xmax = int(np.max(x)) if np.size(x) else 0
    if modulus is None:
        m = int(np.ceil(xmax / size))
        m = max(m, 1)
        while not _nt.is_prime(m * size + 1):
            m += 1
        modulus = m * size + 1
    length = np.shape(x)[-1] if isinstance(x, np.ndarray) else len(x)
    if not size >= length:
        raise ValueError(f"Argument 'size' must be at least the length of the input which is {length}, not {size}.")
    if not _nt.is_prime(int(modulus)):
        raise ValueError(f"Argument 'modulus' must be prime, {modulus} is not.")
    if not (modulus - 1) % size == 0:
        raise ValueError("Argument 'modulus' must equal m * size + 1, where 'size' is the size of the NTT transform.")
    if not modulus > xmax:
        raise ValueError(f"Argument 'modulus' must be at least the max value of the input which is {xmax}, not {modulus}.")
    field = GF(int(modulus))
    x = x if isinstance(x, field) else field(x)
    if forward:
        return np.fft.fft(x, n=int(size))
    norm = "backward" if scaled else "forward"
    return np.fft.ifft(x, n=int(size), norm=norm)


def ntt(x, size=None, modulus=None):
    """Number-theoretic transform X_k = sum_j x_j omega^(jk) (galois.ntt, _ntt.py:16-118)."""
    if not isinstance(x, (tuple, list, np.ndarray, FieldArray)):
        raise TypeError(f"Argument 'x' must be array-like, not {type(x)}.")
    if isinstance(x, FieldArray) and not type(x)._is_prime_field:
        raise ValueError(f"If argument 'x' is a FieldArray, it must be a prime field, not {type(x)}.")
    if modulus is None and isinstance(x, FieldArray):
        modulus = type(x).characteristic
    return _ntt(x, size=size, modulus=modulus, forward=True)


def intt(X, size=None, modulus=None, scaled=True):
    """Inverse NTT (galois.intt, _ntt.py:121-236); ``scaled=False`` skips the division by N."""
    if not isinstance(X, (tuple, list, np.ndarray, FieldArray)):
        raise TypeError(f"Argument 'X' must be array-like, not {type(X)}.")
    if isinstance(X, FieldArray) and not type(X)._is_prime_field:
        raise ValueError(f"If argument 'X' is a FieldArray, it must be a prime field, not {type(X)}.")
    if not isinstance(scaled, bool):
        raise TypeError(f"Argument 'scaled' must be a bool, not {type(scaled)}.")
    if modulus is None and isinstance(X, FieldArray):
        modulus = type(X).characteristic
    return _ntt(X, size=size, modulus=modulus, forward=False, scaled=scaled)


# ---------------------------------------------------------------------------------- device-resident
class _Workspace:
    """Per-device scratch tensor for gfb_ntt_exec, grown on demand."""

    tensors: dict = {}

    @classmethod
    def get(cls, nbytes: int):
        import torch

        dev = torch.cuda.current_device()
        t = cls.tensors.get(dev)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device="cuda")
            cls.tensors[dev] = t
        return t


def ntt_device(x: DeviceArray, inverse: bool = False, scaled: bool | None = None, out: DeviceArray | None = None):
    """
    Batched transform of a device-resident [batch, N] (or [N]) array; rows are independent.
    ``scaled`` defaults to the reference's behaviour (forward: no, inverse: yes).
    """
    field = x.field
    n = x.shape[-1]
    _check_size(field, n)
    if x.elem_bytes not in (4, 8):
        raise TypeError("Device transforms need a 4- or 8-byte dtype.")
    scaled = inverse if scaled is None else scaled
    plan = _plan(field, n, inverse, scaled)
    batch = x.size // n
    lib = _lib.lib()
    out = DeviceArray.empty(field, x.shape, x.dtype) if out is None else out
    ws_bytes = lib.gfb_ntt_workspace_bytes(plan, x.elem_bytes, batch)
    ws = _Workspace.get(ws_bytes) if ws_bytes else None
    st = lib.gfb_ntt_exec(plan, x.elem_bytes, x.data_ptr, out.data_ptr, batch, ws.data_ptr() if ws is not None else None,
                          DeviceArray._stream())
    _lib.check(st, "gfb_ntt_exec")
    return out
