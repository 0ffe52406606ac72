"""
CPU tests of galois_b200.enable(): the hook plumbing under the reference's dispatch with the device calls of the mirror
class STUBBED by the oracle (there is no GPU in the build container; the real thing is tests/test_gpu_binding.py).
What is checked here is everything between the reference's dispatcher and the C-ABI call: mode registration, operand
views, dtype rule, out= / where=, exception parity, fallback of non-accelerated functions to the reference's own JIT code.
"""

import numpy as np
import pytest

from refimport import import_reference

galois = import_reference()
if galois is None:  # pragma: no cover
    pytest.skip("the reference is not installed at baseline/_ref", allow_module_level=True)

import galois_b200 as gb  # noqa: E402
import oracle  # noqa: E402
from galois_b200 import _binding, _lib  # noqa: E402


@pytest.fixture()
def stubbed(monkeypatch):
    """Pretend the CUDA library is there; answer the mirror's device calls with the oracle."""
    monkeypatch.setattr(_lib, "lib", lambda: object())

    def ofield(M):
        return oracle.Field.prime(M.characteristic) if M._is_prime_field else oracle.Field.binary(M.degree, int(M.irreducible_poly))

    def _binary(cls, op, a, b, dtype):
        a, b = np.broadcast_arrays(np.asarray(a), np.asarray(b))
        try:
            z = ofield(cls).binary_op(op, a.astype(np.int64).ravel(), b.astype(np.int64).ravel())
        except ZeroDivisionError:
            raise ZeroDivisionError(_lib.ZERO_DIVISION_MESSAGE) from None
        return z.astype(dtype).reshape(a.shape).view(cls)

    def _unary(cls, op, a, dtype):
        a = np.asarray(a)
        try:
            z = ofield(cls).unary_op(op, a.astype(np.int64).ravel())
        except ZeroDivisionError:
            raise ZeroDivisionError(_lib.ZERO_DIVISION_MESSAGE) from None
        return z.astype(dtype).reshape(a.shape).view(cls)

    def _power(cls, inputs, operands, dtype, outer=False):
        a, e = np.broadcast_arrays(np.asarray(inputs[0]), np.asarray(inputs[1]))
        try:
            z = ofield(cls).power(a.astype(np.int64).ravel(), e.astype(np.int64).ravel())
        except ZeroDivisionError:
            raise ZeroDivisionError(_lib.ZERO_DIVISION_MESSAGE) from None
        return z.astype(dtype).reshape(a.shape).view(cls)

    from galois_b200._field import FieldArray

    monkeypatch.setattr(FieldArray, "_binary", classmethod(_binary))
    monkeypatch.setattr(FieldArray, "_unary", classmethod(_unary))
    monkeypatch.setattr(FieldArray, "_power", classmethod(_power))

    def _fold(cls, inner_op, x, axis, scan, dtype):
        x = np.moveaxis(np.asarray(x).astype(np.int64), axis, -1)
        acc = x[..., 0].copy()
        outs = [acc.copy()]
        for j in range(1, x.shape[-1]):
            acc = ofield(cls).binary_op(inner_op, acc.reshape(-1), x[..., j].reshape(-1)).reshape(acc.shape)
            outs.append(acc.copy())
        res = np.moveaxis(np.stack(outs, axis=-1), -1, axis) if scan else acc
        return res.astype(dtype)

    monkeypatch.setattr(FieldArray, "_fold", classmethod(_fold))

    def convolve(field, a, b, mode="full"):
        c = ofield(field).convolve(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64))
        return c.astype(a.dtype).view(field)

    from galois_b200 import _linalg

    monkeypatch.setattr(_linalg, "convolve", convolve)
    yield


def both(GF, fn):
    gb.disable(GF)
    try:
        want, err = fn(), None
    except Exception as e:  # noqa: BLE001
        want, err = None, e
    gb.enable(GF)
    try:
        if err is not None:
            with pytest.raises(type(err)) as info:
                fn()
            assert str(info.value) == str(err)
            return
        got = fn()
    finally:
        gb.disable(GF)
    assert type(got) is type(want)
    if isinstance(got, np.ndarray):
        assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)


def test_enable_requires_the_backend():
    GF = galois.GF(31)
    with pytest.raises(gb.BackendUnavailable):
        gb.enable(GF)
    assert GF.ufunc_mode != "cuda"


@pytest.mark.parametrize("order", [31, 2**31 - 1, 2**8])
def test_dispatch_plumbing(stubbed, order):
    GF = galois.GF(order)
    prev = GF.ufunc_mode
    gb.enable(GF)
    assert GF.ufunc_mode == "cuda" and GF.ufunc_modes[-1] == "cuda"
    assert isinstance(GF._multiply.ufunc, _binding.CudaUFunc)
    gb.disable(GF)
    assert GF.ufunc_mode == prev

    x = GF.Random(200, seed=1)
    y = GF.Random(200, seed=2, low=1)
    wide = GF.dtypes[-1]
    both(GF, lambda: x * y)
    both(GF, lambda: x / y)
    both(GF, lambda: x - y)
    both(GF, lambda: y**-1)
    both(GF, lambda: x**3)
    both(GF, lambda: np.square(x))
    both(GF, lambda: x * 7)
    both(GF, lambda: x.astype(wide) * y)
    both(GF, lambda: x * y.astype(wide))
    both(GF, lambda: np.multiply.outer(x[:5], y[:7]))
    both(GF, lambda: x[3] * y[4])
    both(GF, lambda: x / GF.Zeros(200))
    both(GF, lambda: GF.Zeros(3) ** -1)
    both(GF, lambda: x + 1)
    both(GF, lambda: x ** GF(2))
    both(GF, lambda: np.negative.reduce(x))
    both(GF, lambda: np.multiply.accumulate(y[:20]))
    both(GF, lambda: np.add.reduce(x))
    both(GF, lambda: np.subtract.accumulate(x[:9]))
    both(GF, lambda: np.add.reduce(x.reshape(10, 20), axis=1))

    def with_out():
        o = GF.Zeros(200, dtype=wide)
        assert np.multiply(x, y, out=o) is o
        return o

    both(GF, with_out)

    def with_where():
        o = GF.Ones(200)
        np.divide(x, GF.Zeros(200) + (np.arange(200) % 2 == 0).astype(GF.dtypes[0]).view(GF), out=o, where=np.arange(200) % 2 == 0)
        return o

    both(GF, with_where)


def test_compile_accepts_cuda_after_enable(stubbed):
    GF = galois.GF(3191)
    with pytest.raises(ValueError):
        GF.compile("cuda")
    gb.enable(GF)
    GF.compile("jit-calculate")
    assert GF.ufunc_mode == "jit-calculate"
    GF.compile("cuda")
    assert GF.ufunc_mode == "cuda"
    GF.compile("auto")
    assert GF.ufunc_mode == GF.default_ufunc_mode
    gb.disable(GF)


def test_non_accelerated_functions_run_the_reference_jit(stubbed):
    """Polynomial algebra has no CUDA implementation: in mode 'cuda' it must run the reference's own compiled function
    (compiled against the class's previous mode), not crash on the mode assertion (_domains/_function.py:86)."""
    GF = galois.GF(3191)
    f = galois.Poly.Random(9, seed=1, field=GF)
    g = galois.Poly.Random(4, seed=2, field=GF)
    both(GF, lambda: divmod(f, g)[0].coeffs)
    both(GF, lambda: divmod(f, g)[1].coeffs)
    both(GF, lambda: galois.gcd(f * g, g * g).coeffs)
    both(GF, lambda: galois.berlekamp_massey(GF.Random(20, seed=3)).coeffs)


def test_unsupported_fields(stubbed):
    with pytest.raises(NotImplementedError):
        gb.enable(galois.GF(3**4))
    with pytest.raises(NotImplementedError):
        gb.enable(galois.GF(36893488147419103183))
    with pytest.raises(TypeError):
        gb.enable(int)
