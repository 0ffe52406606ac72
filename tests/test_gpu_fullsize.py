"""
BASELINE.json's full sizes through size-independent properties (``-m gpu``): where the oracle would take minutes,
the domain's own identities check every element — inverse round trips, linearity, transform round trips and
encode → corrupt → decode round trips — plus an oracle spot check on a slice.
"""

import numpy as np
import pytest

import galois_b200 as gb
import oracle

pytestmark = pytest.mark.gpu


def _dev_eq(a, b) -> bool:
    import torch

    return bool(torch.equal(a.tensor, b.tensor))


@pytest.mark.parametrize("dtype", [np.int64, np.uint32])
def test_config2_gfp_1e8(dtype):
    """GF(2^31-1), 1e8 elements (BASELINE configs[1]): (x*y)/y == x, (x+y)-y == x, x*y == y*x, y*y^-1 == 1,
    distributivity, device-resident; oracle on the first 1e5."""
    import torch

    GF = gb.GF(2**31 - 1)
    n = 10**8
    p = GF.order
    tdt = torch.int64 if dtype == np.int64 else torch.int32
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randint(0, p, (n,), dtype=torch.int64, device="cuda", generator=g).to(tdt)
    y = torch.randint(1, p, (n,), dtype=torch.int64, device="cuda", generator=g).to(tdt)
    z = torch.randint(0, p, (n,), dtype=torch.int64, device="cuda", generator=g).to(tdt)
    dx = gb.DeviceArray(GF, x.view(torch.uint8), (n,), dtype)
    dy = gb.DeviceArray(GF, y.view(torch.uint8), (n,), dtype)
    dz = gb.DeviceArray(GF, z.view(torch.uint8), (n,), dtype)
    xy = dx * dy
    assert _dev_eq(xy / dy, dx)
    assert _dev_eq(xy, dy * dx)
    assert _dev_eq((dx + dy) - dy, dx)
    assert _dev_eq(dx * (dy + dz), xy + dx * dz)
    one = dy * dy.reciprocal()
    assert bool((one.tensor.view(tdt) == 1).all())
    assert _dev_eq(dy ** -1, dy.reciprocal())
    assert _dev_eq(dy ** (p - 2), dy.reciprocal())
    F = oracle.Field.prime(p)
    k = 100_000
    xs, ys = x[:k].cpu().numpy().astype(np.int64), y[:k].cpu().numpy().astype(np.int64)
    assert np.array_equal(xy.tensor.view(tdt)[:k].cpu().numpy().astype(np.int64), F.mul(xs, ys))
    assert np.array_equal(dy.reciprocal().tensor.view(tdt)[:k].cpu().numpy().astype(np.int64), F.inv(ys))
    # a single zero anywhere must raise, exactly like the reference's ZeroDivisionError from inside the loop
    y[n - 3] = 0
    with pytest.raises(ZeroDivisionError):
        dx / dy


def test_config3_ntt_batch():
    """2^20-point NTT over GF(998244353), 256 rows: intt(ntt(x)) == x on every row, linearity, Parseval-like
    convolution identity on a pair of rows; oracle on two rows."""
    import torch

    GF = gb.GF(998244353)
    p, N, rows = GF.order, 2**20, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randint(0, p, (rows, N), dtype=torch.int32, device="cuda", generator=g)
    y = torch.randint(0, p, (rows, N), dtype=torch.int32, device="cuda", generator=g)
    dx = gb.DeviceArray(GF, x.view(torch.uint8).reshape(-1), (rows, N), np.uint32)
    dy = gb.DeviceArray(GF, y.view(torch.uint8).reshape(-1), (rows, N), np.uint32)
    X = gb.ntt_device(dx)
    assert _dev_eq(gb.ntt_device(X, inverse=True), dx)
    Y = gb.ntt_device(dy)
    assert _dev_eq(gb.ntt_device(dx + dy), X + Y)
    # circular convolution theorem on row 0: intt(X*Y) == cyclic convolution; check a few coefficients exactly
    prod = gb.ntt_device(X * Y, inverse=True)
    x0 = x[0].cpu().numpy().astype(object)
    y0 = y[0].cpu().numpy().astype(object)
    got = prod.tensor.view(torch.int32)[:N].cpu().numpy()
    for k in (0, 1, 12345, N - 1):
        idx = (k - np.arange(N)) % N
        assert int(got[k]) == int(np.dot(x0, y0[idx]) % p)
    F = oracle.Field.prime(p)
    omega = int(GF.primitive_root_of_unity(N))
    for r in (0, rows - 1):
        want = F.fft(x[r].cpu().numpy().astype(np.int64), omega)
        assert np.array_equal(X.tensor.view(torch.int32)[r * N:(r + 1) * N].cpu().numpy().astype(np.int64), want)


def test_config4_gf2_128_1e6():
    """GF(2^128) GCM polynomial, 1e6 elements: a*a^-1 == 1, commutativity, distributivity over xor; oracle slice."""
    import torch

    GF = gb.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
    n = 10**6
    g = torch.Generator(device="cuda").manual_seed(5)

    def rand():
        t = torch.randint(-(2**63), 2**63 - 1, (n, 2), dtype=torch.int64, device="cuda", generator=g)
        t[:, 0] |= 1
        return gb.DeviceArray(GF, t.view(torch.uint8).reshape(-1), (n,), np.object_), t

    (a, ta), (b, tb), (c, _) = rand(), rand(), rand()
    ab = a * b
    assert _dev_eq(ab, b * a)
    assert _dev_eq(a * (b + c), ab + a * c)
    assert _dev_eq(ab / b, a)
    one = (a * a.reciprocal()).tensor.view(torch.int64).view(n, 2)
    assert bool((one[:, 0] == 1).all()) and bool((one[:, 1] == 0).all())
    assert _dev_eq(a.square(), a * a)
    O = oracle.Field128(128, (1 << 128) | 0x87)
    k = 20000
    la = ta[:k].cpu().numpy().view(np.uint64)
    lb = tb[:k].cpu().numpy().view(np.uint64)
    assert np.array_equal(ab.tensor[: 16 * k].cpu().numpy().view(np.uint64).reshape(-1, 2), O.mul(la, lb, threads=8))
    assert np.array_equal(a.reciprocal().tensor[: 16 * 200].cpu().numpy().view(np.uint64).reshape(-1, 2), O.inv(la[:200]))


def test_config5_rs_1e6_roundtrip():
    """RS(255,223), 1e6 codewords: encode, 16 random symbol errors each, decode == encoded and n_errors == 16;
    32 erasures likewise; 32 errors -> every row -1 and unchanged."""
    import torch

    rs = gb.ReedSolomon(255, 223)
    GF = rs.field
    N = 10**6
    M = GF.Random((N, 223), seed=6)
    dC = rs.encode_device(M.to_device())
    C = dC.tensor.view(N, 255)
    assert bool((C[:, :223].cpu() == torch.from_numpy(np.asarray(M))).all())
    g = torch.Generator(device="cuda").manual_seed(7)

    def corrupt(n_err, n_era):
        tot = n_err + n_era
        pos = torch.rand((N, 255), device="cuda", generator=g).topk(tot, dim=1).indices
        err = torch.randint(1, 256, (N, tot), device="cuda", generator=g, dtype=torch.int32).to(torch.uint8)
        R = C.clone()
        R.scatter_(1, pos, R.gather(1, pos) ^ err)
        E = None
        if n_era:
            E = torch.zeros((N, 255), dtype=torch.uint8, device="cuda")
            E.scatter_(1, pos[:, n_err:], 1)
        return gb.DeviceArray(GF, R.reshape(-1), (N, 255), np.uint8), E

    for n_err, n_era in [(16, 0), (0, 32), (5, 22)]:
        R, E = corrupt(n_err, n_era)
        dec, nerr = rs.decode_device(R, erasures=E)
        assert bool((nerr == n_err).all()), (n_err, n_era)
        assert bool((dec.tensor.view(N, 255) == C).all()), (n_err, n_era)
    R, _ = corrupt(32, 0)
    dec, nerr = rs.decode_device(R)
    # beyond t: almost always detected as uncorrectable; a detected row is returned unchanged (_bch.py:1344)
    bad = nerr == -1
    assert float(bad.float().mean()) > 0.99
    assert bool((dec.tensor.view(N, 255)[bad] == R.tensor.view(N, 255)[bad]).all())


def test_config1_gf256_1e6_plumbing():
    """GF(2^8) 1e6-element host arrays through __array_ufunc__ (BASELINE configs[0]), every op vs the oracle."""
    GF = gb.GF(2**8)
    F = oracle.Field.binary(8, 0x11D, alpha=2)
    x = GF.Random(10**6, seed=1)
    y = GF.Random(10**6, seed=2, low=1)
    xs, ys = np.asarray(x).astype(np.int64), np.asarray(y).astype(np.int64)
    for got, want in [(x * y, F.mul(xs, ys)), (x + y, F.add(xs, ys)), (x / y, F.div(xs, ys)),
                      (np.reciprocal(y), F.inv(ys)), (x**3, F.power(xs, 3)), (y**-1, F.inv(ys))]:
        assert type(got) is GF and got.dtype == np.uint8
        assert np.array_equal(np.asarray(got), want)


def test_edge_shapes():
    """Empty, 0-d, 1-element, non-contiguous and broadcast operands."""
    for GF in (gb.GF(2**31 - 1), gb.GF(2**8), gb.GF(2**32)):
        e = GF([])
        assert (e * e).shape == (0,) and type(e + e) is GF
        assert np.reciprocal(e).shape == (0,)
        s = GF(5)
        assert s.ndim == 0 and int(s * s) == int(np.asarray(GF([5]) * GF([5]))[0])
        one = GF([3])
        assert int((one / one)[0]) == 1
        A = GF.Random((64, 48), seed=9)
        At = A.T  # non-contiguous view
        assert np.array_equal(np.asarray(At * At), np.asarray((A * A)).T)
        col = A[:, 3]
        assert np.array_equal(np.asarray(col + col), np.asarray(A + A)[:, 3])
        assert np.array_equal(np.asarray(A * A[0]), np.asarray(A * np.broadcast_to(A[0], A.shape).view(GF)))
    GN = gb.GF(998244353)
    assert np.array_equal(np.asarray(gb.ntt(GN([7]))), [7])
    assert np.array_equal(np.asarray(gb.ntt(GN([1, 2]))), [3, 998244352])
