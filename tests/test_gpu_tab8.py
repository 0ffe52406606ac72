"""
GPU tests of the product-table kernels for byte-sized binary fields (gfb_ewise_tab8.cu, ``-m gpu``): every GF(2^m),
m <= 8, multiply and divide, array-array and both scalar modes, ragged sizes, zero divisors — against the oracle's
restatement of the reference's lookup ufuncs (_domains/_lookup.py:143-233).  The reference's full 256 x 256 Sage tables
for GF(2^8) run through the same kernels in tests/test_gpu_parity.py::test_ref_vectors_arithmetic.
"""

import numpy as np
import pytest

import galois_b200 as gb
import oracle

pytestmark = pytest.mark.gpu

POLYS = {1: 0b11, 2: 0b111, 3: 0b1011, 4: 0b10011, 5: 0b100101, 6: 0b1011011, 7: 0b10000011, 8: 0b100011101}


def _field(m):
    if m == 1:
        return gb.GF(2), oracle.Field.prime(2)
    GF = gb.GF(2**m, irreducible_poly=POLYS[m])
    return GF, oracle.Field.binary(m, POLYS[m], alpha=int(GF.primitive_element))


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5, 6, 7, 8])
def test_full_tables(m):
    GF, F = _field(m)
    q = 2**m
    A, B = np.meshgrid(np.arange(q), np.arange(q), indexing="ij")
    a, b = GF(A.ravel().astype(np.uint8)), GF(B.ravel().astype(np.uint8))
    assert np.array_equal(np.asarray(a * b), F.mul(A.ravel(), B.ravel()))
    nz = B.ravel() != 0
    assert np.array_equal(np.asarray(a[nz] / b[nz]), F.div(A.ravel()[nz], B.ravel()[nz]))
    with pytest.raises(ZeroDivisionError):
        a / b


@pytest.mark.parametrize("n", [1, 15, 16, 17, 31, 4096, 1_000_003])
def test_ragged_sizes_and_scalars(n):
    GF, F = _field(8)
    rng = np.random.default_rng(n)
    x = rng.integers(0, 256, n, dtype=np.uint8)
    y = rng.integers(1, 256, n, dtype=np.uint8)
    gx, gy = GF(x), GF(y)
    assert np.array_equal(np.asarray(gx * gy), F.mul(x, y))
    assert np.array_equal(np.asarray(gx / gy), F.div(x, y))
    s = 0x53
    assert np.array_equal(np.asarray(gx * GF(s)), F.mul(x, np.full(n, s)))
    assert np.array_equal(np.asarray(GF(s) * gx), F.mul(x, np.full(n, s)))
    assert np.array_equal(np.asarray(gx / GF(s)), F.div(x, np.full(n, s)))
    assert np.array_equal(np.asarray(GF(s) / gy), F.div(np.full(n, s), y))
    assert np.array_equal(np.asarray(gx * GF(0)), np.zeros(n, dtype=np.uint8))
    with pytest.raises(ZeroDivisionError):
        gx / GF(0)
    if n > 3:
        y0 = y.copy()
        y0[n - 2] = 0
        with pytest.raises(ZeroDivisionError):
            gx / GF(y0)
        with pytest.raises(ZeroDivisionError):
            GF(s) / GF(y0)
    # wider storage of the same field keeps the EXP/LOG kernels: both paths agree
    assert np.array_equal(np.asarray(GF(x, dtype=np.int32) * GF(y, dtype=np.int32)), F.mul(x, y))


def test_device_resident_1e8():
    import torch

    GF, F = _field(8)
    n = 10**8
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(1, 256, (n,), dtype=torch.uint8, device="cuda", generator=g)
    da = gb.DeviceArray(GF, a, (n,), np.uint8)
    db = gb.DeviceArray(GF, b, (n,), np.uint8)
    prod = da * db
    quot = prod / db
    assert torch.equal(quot.tensor, a)
    k = 200_000
    assert np.array_equal(prod.tensor[:k].cpu().numpy(), F.mul(a[:k].cpu().numpy(), b[:k].cpu().numpy()))
    assert np.array_equal(prod.tensor[-k:].cpu().numpy(), F.mul(a[-k:].cpu().numpy(), b[-k:].cpu().numpy()))
