"""
GPU tests of the LIVE binding (``-m gpu``): ``galois_b200.enable(GF)`` on the unmodified reference's own field classes
(installed at baseline/_ref by the recipe in DESIGN.md; it travels to the GPU box with the snapshot).  Every case runs the
same reference call twice on identical inputs — once in the class's stock mode (Numba / pure Python), once in mode
"cuda" — and requires bit-equal values, identical result type and dtype, and identical exceptions.
"""

import numpy as np
import pytest

from refimport import import_reference

pytestmark = pytest.mark.gpu

galois = import_reference()
if galois is None:  # pragma: no cover
    pytest.skip("the reference is not installed at baseline/_ref", allow_module_level=True)

import galois_b200 as gb  # noqa: E402


def same(a, b):
    if isinstance(a, tuple):
        assert isinstance(b, tuple) and len(a) == len(b)
        for x, y in zip(a, b):
            same(x, y)
        return
    assert type(a) is type(b), (type(a), type(b))
    if isinstance(a, np.ndarray):
        assert a.dtype == b.dtype, (a.dtype, b.dtype)
        assert a.shape == b.shape, (a.shape, b.shape)
        assert np.array_equal(np.asarray(a).astype(object), np.asarray(b).astype(object))
    else:
        assert a == b


def both(fields, fn):
    """fn() under the stock mode(s) and under mode 'cuda'; returns the cuda result after comparing."""
    fields = fields if isinstance(fields, (list, tuple)) else [fields]
    for f in fields:
        gb.disable(f)
    try:
        want = fn()
        err = None
    except Exception as e:  # noqa: BLE001
        want, err = None, e
    for f in fields:
        gb.enable(f)
        assert f.ufunc_mode == "cuda"
    try:
        if err is not None:
            with pytest.raises(type(err)) as info:
                fn()
            assert str(info.value) == str(err)
            return None
        got = fn()
    finally:
        for f in fields:
            gb.disable(f)
    same(got, want)
    return got


def test_enable_is_a_mode_of_the_reference_class():
    GF = galois.GF(2**31 - 1)
    prev = GF.ufunc_mode
    assert gb.enable(GF) is GF
    assert GF.ufunc_mode == "cuda" and "cuda" in GF.ufunc_modes
    x = GF.Random(10, seed=1)
    z = x * x
    assert type(z) is GF and isinstance(z, galois.FieldArray) and isinstance(z, np.ndarray)
    GF.compile(prev)
    assert GF.ufunc_mode == prev
    GF.compile("cuda")
    assert GF.ufunc_mode == "cuda"
    GF.compile("auto")
    assert GF.ufunc_mode == GF.default_ufunc_mode
    gb.disable(GF)
    with pytest.raises(TypeError):
        gb.enable(np.ndarray)
    with pytest.raises(NotImplementedError):
        gb.enable(galois.GF(3**5))


@pytest.mark.parametrize("order", [2**31 - 1, 998244353, 31, 3191, 2**8, 2**4, 2**16, 2**32])
def test_elementwise_ufuncs(order):
    GF = galois.GF(order)
    for dtype in GF.dtypes[:1] + GF.dtypes[-1:]:
        x = GF.Random(5000, seed=1, dtype=dtype)
        y = GF.Random(5000, seed=2, low=1, dtype=dtype)
        e = np.random.default_rng(3).integers(-50, 50, 5000)
        edge = GF([0, 1, 2 % order, order - 2, order - 1], dtype=dtype)
        both(GF, lambda: x * y)
        both(GF, lambda: x + y)
        both(GF, lambda: x - y)
        both(GF, lambda: x / y)
        both(GF, lambda: x // y)
        both(GF, lambda: -x)
        both(GF, lambda: np.reciprocal(y))
        both(GF, lambda: y**-1)
        both(GF, lambda: x**3)
        both(GF, lambda: np.square(x))
        both(GF, lambda: y**e)
        both(GF, lambda: x * 5)
        both(GF, lambda: 7 * x)
        both(GF, lambda: x * np.arange(5000))
        both(GF, lambda: np.multiply.outer(edge, edge))
        both(GF, lambda: np.add.outer(edge, edge))
        both(GF, lambda: edge[:, None] * edge[None, :])
        both(GF, lambda: x[0] * y[0])  # 0-d field scalars
        both(GF, lambda: np.add.reduce(x))
        both(GF, lambda: np.multiply.reduce(y))
        both(GF, lambda: np.add.reduce(x.reshape(50, 100), axis=1))
        both(GF, lambda: np.multiply.accumulate(y[:64]))
        both(GF, lambda: np.add.reduceat(x, [0, 10, 10, 4000]))
        both(GF, lambda: np.divmod(x, y))
        both(GF, lambda: np.remainder(x, y))


def test_mixed_dtypes_and_out():
    GF = galois.GF(2**31 - 1)
    a = GF.Random(1000, seed=1, dtype=np.uint32)
    b = GF.Random(1000, seed=2, dtype=np.int64)
    both(GF, lambda: a * b)
    both(GF, lambda: b * a)
    both(GF, lambda: a + b)

    def with_out():
        o = GF.Zeros(1000, dtype=np.int64)
        r = np.multiply(b, b, out=o)
        assert r is o
        return o

    both(GF, with_out)

    def with_where():
        o = GF.Ones(1000, dtype=np.int64)
        mask = np.arange(1000) % 3 == 0
        np.multiply(b, b, out=o, where=mask)
        return o

    both(GF, with_where)

    def inplace():
        c = a.copy()
        c *= a
        c += a
        c /= GF(3, dtype=np.uint32)
        return c

    both(GF, inplace)


@pytest.mark.parametrize("order", [2**31 - 1, 2**8, 2**32])
def test_zero_division_and_type_errors(order):
    GF = galois.GF(order)
    x = GF.Random(100, seed=1)
    z = GF.Zeros(100)
    both(GF, lambda: x / z)
    both(GF, lambda: np.reciprocal(z))
    both(GF, lambda: z**-1)
    both(GF, lambda: z**-3)
    both(GF, lambda: x + 1)
    both(GF, lambda: x * 1.5)
    both(GF, lambda: x ** GF(2))
    both(GF, lambda: x ** 1.5)
    both(GF, lambda: np.add.reduce(x, axis=5) if False else np.negative.reduce(x))
    both(GF, lambda: np.power.reduce(x))
    both(GF, lambda: np.exp(x))


def test_gf2_128_object_dtype():
    GF = galois.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
    rng = np.random.default_rng(5)
    vals = [int(rng.integers(0, 2**63)) | (int(rng.integers(0, 2**63)) << 64) | 1 for _ in range(64)]
    x = GF(vals[:32])
    y = GF(vals[32:])
    both(GF, lambda: x * y)
    both(GF, lambda: x + y)
    both(GF, lambda: x / y)
    both(GF, lambda: y[:4] ** -1)
    both(GF, lambda: x[:4] ** 5)
    both(GF, lambda: GF([0, 1]) ** -1)


@pytest.mark.parametrize("n", [1, 2, 8, 1024, 2**14, 2**16])
def test_ntt_intt(n):
    p = 998244353
    GF = galois.GF(p)
    x = GF.Random(n, seed=n)
    both(GF, lambda: galois.ntt(x))
    both(GF, lambda: galois.intt(galois.ntt(x)))
    both(GF, lambda: galois.intt(x, scaled=False))
    both(GF, lambda: np.fft.fft(x))
    both(GF, lambda: np.fft.ifft(x))
    if n >= 8:
        both(GF, lambda: np.fft.fft(x, n=n // 2))
        both(GF, lambda: np.fft.fft(x[: n // 2], n=n))
        both(GF, lambda: galois.ntt(list(range(n // 2)), size=n, modulus=p))
    got = both(GF, lambda: galois.ntt(x, modulus=p))
    assert type(got) is GF


def test_ntt_large_and_errors():
    GF = galois.GF(998244353)
    x = GF.Random(2**20, seed=3)
    gb.enable(GF)
    try:
        X = galois.ntt(x)
        assert type(X) is GF
        assert np.array_equal(galois.intt(X), x)
    finally:
        gb.disable(GF)
    # spot check against the reference on a length the CPU finishes quickly
    y = GF.Random(2**17, seed=4)
    both(GF, lambda: galois.ntt(y))
    G31 = galois.GF(31)
    both(G31, lambda: np.fft.fft(G31.Random(7, seed=1)))  # 7 does not divide 30
    both(G31, lambda: np.fft.fft(G31.Random(6, seed=1)))  # mixed radix 2 * 3
    both(G31, lambda: np.fft.fft(G31.Random(6, seed=1), norm="ortho"))
    G256 = galois.GF(2**8)
    both(G256, lambda: np.fft.fft(G256.Random(15, seed=1)))
    both(G256, lambda: np.fft.ifft(G256.Random(85, seed=1)))


@pytest.mark.parametrize("order", [998244353, 2**31 - 1, 2**8, 31])
def test_convolve_matmul(order):
    GF = galois.GF(order)
    a = GF.Random(300, seed=1)
    b = GF.Random(77, seed=2)
    both(GF, lambda: np.convolve(a, b))
    both(GF, lambda: np.convolve(b, a))
    both(GF, lambda: np.convolve(a, b, mode="same"))
    A = GF.Random((40, 33), seed=3)
    B = GF.Random((33, 21), seed=4)
    both(GF, lambda: A @ B)
    both(GF, lambda: np.matmul(A, B))
    both(GF, lambda: A @ B[:, 0])
    both(GF, lambda: A[0, :] @ B)
    both(GF, lambda: A[0, :] @ B[:, 0])
    both(GF, lambda: A @ A)


@pytest.mark.parametrize("order", [998244353, 2**8, 31, 2**16])
def test_linalg(order):
    GF = galois.GF(order)
    A = GF.Random((24, 24), seed=5)
    bvec = GF.Random(24, seed=6)
    W = GF.Random((10, 17), seed=7)
    S = A.copy()
    S[5, :] = S[2, :]  # singular
    both(GF, lambda: A.row_reduce())
    both(GF, lambda: W.row_reduce())
    both(GF, lambda: W.row_reduce(ncols=5))
    both(GF, lambda: np.linalg.inv(A))
    both(GF, lambda: np.linalg.inv(S))
    both(GF, lambda: np.linalg.det(A))
    both(GF, lambda: np.linalg.det(S))
    both(GF, lambda: np.linalg.solve(A, bvec))
    both(GF, lambda: np.linalg.matrix_rank(S))
    both(GF, lambda: A.plu_decompose())
    both(GF, lambda: W.plu_decompose())
    both(GF, lambda: W.null_space())
    both(GF, lambda: W.T.copy().left_null_space())
    both(GF, lambda: S.row_space())


def test_reed_solomon_255_223():
    rs = galois.ReedSolomon(255, 223)
    GF = rs.field
    rng = np.random.default_rng(7)
    N = 300
    M = GF.Random((N, 223), seed=6)
    C = both(GF, lambda: rs.encode(M))
    R = C.copy()
    n_err = rng.integers(0, 18, N)  # up to 17 > t = 16: some rows are uncorrectable
    for i in range(N):
        pos = rng.choice(255, int(n_err[i]), replace=False)
        R[i, pos] ^= GF(rng.integers(1, 256, pos.size, dtype=np.uint8))
    both(GF, lambda: rs.decode(R))
    both(GF, lambda: rs.decode(R, errors=True))
    both(GF, lambda: rs.decode(R, output="codeword", errors=True))
    both(GF, lambda: rs.decode(R[0]))
    both(GF, lambda: rs.decode(C, errors=True))  # clean codewords
    era = np.zeros((N, 255), dtype=bool)
    for i in range(N):
        era[i, rng.choice(255, 20, replace=False)] = True
    R2 = C.copy()
    R2[era] = 0
    for i in range(N):
        pos = rng.choice(255, 5, replace=False)
        R2[i, pos] ^= GF(rng.integers(1, 256, 5, dtype=np.uint8))
    both(GF, lambda: rs.decode(R2, erasures=era, errors=True))
    # shortened
    both(GF, lambda: rs.decode(rs.encode(M[:10, :100]), errors=True))
    both(GF, lambda: rs.detect(R))


def test_rs_other_codes_and_bch():
    for args in [(15, 9), (63, 55), (255, 239)]:
        rs = galois.ReedSolomon(*args)
        GF = rs.field
        M = GF.Random((50, rs.k), seed=1)
        C = both(GF, lambda: rs.encode(M))
        R = C.copy()
        rng = np.random.default_rng(2)
        for i in range(50):
            pos = rng.choice(rs.n, rs.t, replace=False)
            R[i, pos] += GF.Random(rs.t, low=1, seed=int(i))
        both(GF, lambda: rs.decode(R, errors=True))
    bch = galois.BCH(15, 7)
    ext = bch.extension_field
    m = galois.GF2.Random((40, 7), seed=3)
    c = bch.encode(m)
    r = c.copy()
    rng = np.random.default_rng(4)
    for i in range(40):
        pos = rng.choice(15, int(rng.integers(0, 3)), replace=False)
        r[i, pos] ^= galois.GF2(1)
    both(ext, lambda: bch.decode(r, errors=True))


def test_poly_evaluate_and_roots():
    GF = galois.GF(2**8)
    f = galois.Poly.Random(12, seed=1, field=GF)
    x = GF.Random(1000, seed=2)
    both(GF, lambda: f(x))
    both(GF, lambda: f(GF(3)))
    g = galois.Poly.Roots([GF(5), GF(17), GF(200)], field=GF) * galois.Poly([1, 2, 3], field=GF)
    both(GF, lambda: g.roots())
    GP = galois.GF(3191)
    h = galois.Poly.Roots([1, 5, 77, 3000], field=GP)
    both(GP, lambda: h.roots())
    both(GP, lambda: h(GP.Random(100, seed=3)))
    # polynomial algebra keeps running the reference's own JIT functions inside mode "cuda"
    both(GP, lambda: (h * h + h).coeffs)
    both(GP, lambda: divmod(h * h, galois.Poly([1, 2, 3], field=GP))[1].coeffs)


def test_fullsize_multiply_through_the_reference_class():
    """BASELINE configs[1] through the reference's own x * y (10^7 here; the 10^8 case is bench.py's e2e leg)."""
    GF = galois.GF(2**31 - 1)
    x = GF.Random(10**7, seed=1, dtype=np.int64)
    y = GF.Random(10**7, seed=2, low=1, dtype=np.int64)
    got = both(GF, lambda: x * y)
    assert got.dtype == np.int64
    both(GF, lambda: x / y)
