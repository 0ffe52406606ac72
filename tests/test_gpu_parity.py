"""
GPU parity tests (``-m gpu``): the CUDA path, reached through the NumPy-facing mirror and hence the
C ABI, against (a) the committed golden fixtures (reference's own vectors + live-reference outputs)
and (b) the oracle on seeded inputs.  Bit-exact everywhere (integer work).
"""

import hashlib

import numpy as np
import pytest

import galois_b200 as gb
import oracle
from helpers import limbs_to_obj, oracle_field, parse_field

pytestmark = pytest.mark.gpu

REF_FIELDS = ["GF(2)", "GF(2^2)", "GF(2^3)", "GF(2^8)", "GF(2^32)", "GF(2^100)", "GF(5)", "GF(7)", "GF(31)", "GF(3191)",
              "GF(2147483647)"]
NARROW = [f for f in REF_FIELDS if f != "GF(2^100)"]


def make_gf(name):
    spec = parse_field(name)
    if spec["p"] != 2:
        return gb.GF(spec["p"])
    return gb.GF(2 ** spec["m"], irreducible_poly=spec["poly"], primitive_element=spec["alpha"])


def eq(a, b):
    return np.array_equal(np.asarray(a).astype(object), np.asarray(b).astype(object))


@pytest.mark.parametrize("name", NARROW)
def test_ref_vectors_arithmetic(ref_fields, name):
    """The reference's Sage vectors (tests/fields/test_arithmetic.py) through the CUDA ufuncs, every dtype."""
    data, _ = ref_fields
    GF = make_gf(name)
    for dtype in GF.dtypes:
        for op, fn in [("add", np.add), ("subtract", np.subtract), ("multiply", np.multiply), ("divide", np.true_divide)]:
            X, Y = np.meshgrid(data[f"{name}/{op}/X"], data[f"{name}/{op}/Y"], indexing="ij")
            x, y = GF(X, dtype=dtype), GF(Y, dtype=dtype)
            z = fn(x, y)
            assert type(z) is GF and z.dtype == dtype
            assert eq(z, data[f"{name}/{op}/Z"]), (name, op, dtype)
        x = GF(data[f"{name}/additive_inverse/X"], dtype=dtype)
        assert eq(-x, data[f"{name}/additive_inverse/Z"])
        x = GF(data[f"{name}/multiplicative_inverse/X"], dtype=dtype)
        z = np.reciprocal(x)
        assert type(z) is GF and z.dtype == dtype
        assert eq(z, data[f"{name}/multiplicative_inverse/Z"])
        assert eq(x ** -1, data[f"{name}/multiplicative_inverse/Z"])
        X, Y = np.meshgrid(data[f"{name}/power/X"], data[f"{name}/power/Y"], indexing="ij")
        z = GF(X, dtype=dtype) ** Y
        assert type(z) is GF and z.dtype == dtype
        assert eq(z, data[f"{name}/power/Z"]), (name, "power", dtype)


def test_ref_vectors_gf2_100(ref_fields):
    data, meta = ref_fields
    name = "GF(2^100)"
    GF = gb.GF(2**100, irreducible_poly=meta[name]["irreducible_poly"], primitive_element=2, verify=False)
    for op, fn in [("add", np.add), ("subtract", np.subtract), ("multiply", np.multiply), ("divide", np.true_divide)]:
        X = limbs_to_obj(data[f"{name}/{op}/X"])
        Y = limbs_to_obj(data[f"{name}/{op}/Y"])
        Z = limbs_to_obj(data[f"{name}/{op}/Z"])
        xx, yy = np.meshgrid(X, Y, indexing="ij")
        z = fn(GF(xx), GF(yy))
        assert type(z) is GF and z.dtype == np.object_
        assert eq(z, Z), op
    X = limbs_to_obj(data[f"{name}/multiplicative_inverse/X"])
    assert eq(np.reciprocal(GF(X)), limbs_to_obj(data[f"{name}/multiplicative_inverse/Z"]))
    X = limbs_to_obj(data[f"{name}/power/X"])
    Y = data[f"{name}/power/Y"]
    xx, ee = np.meshgrid(X, Y, indexing="ij")
    assert eq(GF(xx) ** ee, limbs_to_obj(data[f"{name}/power/Z"]))


@pytest.mark.parametrize("name", NARROW)
def test_ref_vectors_convolve_matmul(ref_fields, name):
    data, _ = ref_fields
    GF = make_gf(name)
    for i in range(int(data[f"{name}/convolve/count"])):
        z = np.convolve(GF(data[f"{name}/convolve/X{i}"]), GF(data[f"{name}/convolve/Y{i}"]))
        assert type(z) is GF
        assert eq(z, data[f"{name}/convolve/Z{i}"])
    for i in range(int(data[f"{name}/matrix_multiply/count"])):
        z = GF(data[f"{name}/matrix_multiply/X{i}"]) @ GF(data[f"{name}/matrix_multiply/Y{i}"])
        assert eq(z, data[f"{name}/matrix_multiply/Z{i}"])


def test_live_elementwise(live_elementwise):
    data, _ = live_elementwise
    keys = sorted({k.rsplit("/", 1)[0] for k in data.files})
    for key in keys:
        name, dtname = key.split("/")
        GF = make_gf(name)
        dtype = np.dtype(dtname)
        g = lambda s: data[f"{key}/{s}"]  # noqa: E731
        x, y, ynz = GF(g("x"), dtype=dtype), GF(g("y"), dtype=dtype), GF(g("ynz"), dtype=dtype)
        for got, want in [(x + y, "add"), (x - y, "sub"), (x * y, "mul"), (x / ynz, "div"), (-x, "neg"),
                          (np.reciprocal(ynz), "inv"), (x**2, "square"), (np.square(x), "square"),
                          (GF(g("xe"), dtype=dtype) ** g("e"), "pow"), (x**3, "pow3"), (ynz**-1, "powm1"),
                          (x * GF(7 % GF.order), "mul_scalar7"), (GF(5 % GF.order) - x, "scalar5_sub"),
                          (GF(3 % GF.order) / ynz, "scalar3_div")]:
            assert type(got) is GF, (key, want)
            # result dtype = dtype of the operand whose __array_ufunc__ ran: the array's, or the 0-d field scalar's
            # (default) dtype when the scalar is on the left (verified on the reference)
            expect = np.dtype(GF.dtypes[0]) if want in ("scalar5_sub", "scalar3_div") else dtype
            assert got.dtype == expect, (key, want)
            assert np.array_equal(np.asarray(got), g(want)), (key, want)
        # device-resident path (no PCIe per op)
        dx, dy, dnz = x.to_device(), y.to_device(), ynz.to_device()
        assert np.array_equal(np.asarray((dx * dy).cpu()), g("mul")), key
        assert np.array_equal(np.asarray((dx + dy).cpu()), g("add")), key
        assert np.array_equal(np.asarray((dx - dy).cpu()), g("sub")), key
        assert np.array_equal(np.asarray((dx / dnz).cpu()), g("div")), key
        assert np.array_equal(np.asarray(dnz.reciprocal().cpu()), g("inv")), key
        assert np.array_equal(np.asarray((dx ** 3).cpu()), g("pow3")), key


def test_zero_division_and_type_errors():
    """Exception parity: tests/fields/test_arithmetic_exceptions.py."""
    for GF in [gb.GF(2**31 - 1), gb.GF(2**8), gb.GF(2**32), gb.GF(31)]:
        x = GF.Random(100, seed=1)
        y = GF.Random(100, seed=2)
        y[7] = 0
        with pytest.raises(ZeroDivisionError, match="Cannot compute the multiplicative inverse of 0 in a Galois field."):
            x / y
        with pytest.raises(ZeroDivisionError):
            np.reciprocal(y)
        with pytest.raises(ZeroDivisionError):
            y ** -1
        with pytest.raises(ZeroDivisionError):
            y.to_device().reciprocal()
        assert np.array_equal(np.asarray(y**0), np.ones(100, dtype=y.dtype))
        with pytest.raises(TypeError):
            x + 1
        with pytest.raises(TypeError):
            x - np.ones(100, dtype=np.int64)
        with pytest.raises(TypeError):
            x / 2
        with pytest.raises(TypeError):
            x ** GF(2)
        with pytest.raises(TypeError):
            GF(np.array([1.0, 2.0]))
        with pytest.raises(ValueError):
            GF([GF.order])
        z = x * 3  # scalar multiplication by an integer is allowed (_ufunc.py:392-401)
        assert type(z) is GF and np.array_equal(np.asarray(z), np.asarray(x * GF(3 % GF.characteristic)))


def test_broadcast_reduce_outer_and_tails():
    """Broadcasting, ufunc methods (tests/fields/test_numpy_ufuncs.py) and odd sizes / unaligned views."""
    GF = gb.GF(2**31 - 1)
    F = oracle.Field.prime(2**31 - 1)
    rng = np.random.default_rng(9)
    for n in [0, 1, 2, 3, 15, 16, 17, 255, 1023, 4097, 100003]:
        a = rng.integers(0, GF.order, n)
        b = rng.integers(1, GF.order, n)
        for dtype in (np.uint32, np.int64):
            x, y = GF(a, dtype=dtype), GF(b, dtype=dtype)
            assert np.array_equal(np.asarray(x * y), F.mul(a, b))
            assert np.array_equal(np.asarray(x / y), F.div(a, b))
            if n > 3:  # unaligned (offset) views exercise the scalar-load fallback kernels
                assert np.array_equal(np.asarray(x[1:] * y[1:]), F.mul(a[1:], b[1:]))
                assert np.array_equal(np.asarray(x[::2] + y[::2]), F.add(a[::2], b[::2]))
    A = GF.Random((7, 1, 5), seed=3)
    B = GF.Random((4, 5), seed=4)
    assert np.array_equal(np.asarray(A * B), F.mul(np.asarray(A), np.asarray(B)))
    assert np.array_equal(np.asarray(np.multiply.outer(A[:, 0, 0], B[0])),
                          F.mul(np.asarray(A[:, 0, 0])[:, None], np.asarray(B[0])[None, :]))
    v = GF.Random(1000, seed=5, low=1)
    total = int(np.add.reduce(v))
    assert total == sum(int(t) for t in v) % GF.order
    prod = 1
    for t in v:
        prod = prod * int(t) % GF.order
    assert int(np.multiply.reduce(v)) == prod
    M = GF.Random((6, 9), seed=6)
    assert np.array_equal(np.asarray(np.add.reduce(M, axis=1)), np.asarray(M).astype(object).sum(axis=1) % GF.order)
    out = GF.Zeros(1000)
    np.multiply(v, v, out=out)
    assert np.array_equal(np.asarray(out), F.mul(np.asarray(v), np.asarray(v)))


def test_live_gf2_128(live_gf128):
    d = live_gf128
    GF = gb.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
    a, b = GF.from_limbs(d["a"]), GF.from_limbs(d["b"])
    assert eq(a * b, limbs_to_obj(d["mul"]))
    assert eq(a + b, limbs_to_obj(d["add"]))
    assert eq(a**2, limbs_to_obj(d["square"]))
    bi = GF.from_limbs(d["inv_in"])
    assert eq(np.reciprocal(bi), limbs_to_obj(d["inv"]))
    assert eq(a[: bi.size] / bi, limbs_to_obj(d["div"]))
    assert eq(GF.from_limbs(d["pow_in"]) ** d["pow_e"], limbs_to_obj(d["pow"]))
    z = a * b
    assert type(z) is GF and z.dtype == np.object_
    # algebra on a larger seeded sample, against the oracle
    rng = np.random.default_rng(55)
    la = rng.integers(0, 2**64, (5000, 2), dtype=np.uint64)
    lb = rng.integers(0, 2**64, (5000, 2), dtype=np.uint64)
    O = oracle.Field128(128, (1 << 128) | 0x87)
    assert np.array_equal((GF.from_limbs(la) * GF.from_limbs(lb)).to_limbs(), O.mul(la, lb))
    inv = np.reciprocal(GF.from_limbs(la[:300]))
    assert all(int(v) == 1 for v in inv * GF.from_limbs(la[:300]))
    poly100 = int(d["gf100_poly"][0, 0]) | (int(d["gf100_poly"][0, 1]) << 64)
    GF100 = gb.GF(2**100, irreducible_poly=poly100, primitive_element=2, verify=False)
    assert eq(GF100.from_limbs(d["gf100_a"]) * GF100.from_limbs(d["gf100_b"]), limbs_to_obj(d["gf100_mul"]))
    assert eq(np.reciprocal(GF100.from_limbs(d["gf100_inv_in"])), limbs_to_obj(d["gf100_inv"]))


# SymPy LUTs, /root/reference/tests/fields/test_ntt.py:13-18
NTT_LUTS = [([1, 2, 3, 4], 5, [0, 4, 3, 2]), ([1, 2, 3, 4], 13, [10, 8, 11, 1]), ([1, 2, 3, 4], 17, [10, 6, 15, 7]),
            ([1, 2, 3, 4], 3 * 256 + 1, [10, 643, 767, 122])]


@pytest.mark.parametrize(["x", "p", "X"], NTT_LUTS)
def test_ntt_sympy(x, p, X):
    GF = gb.GF(p)
    for arg in (tuple(x), list(x), np.array(x), GF(x)):
        got = gb.ntt(arg, modulus=p)
        assert isinstance(got, GF) and np.array_equal(np.asarray(got), X)
    assert np.array_equal(np.asarray(gb.intt(gb.ntt(x, modulus=p))), x)


def test_ntt_exceptions():
    with pytest.raises(TypeError):
        gb.ntt(1)
    with pytest.raises(TypeError):
        gb.ntt([1, 2, 3, 4], size=6.0)
    with pytest.raises(ValueError):
        gb.ntt(gb.GF(2**8)([1, 2, 3, 4]))
    with pytest.raises(ValueError):
        gb.ntt([1, 2, 3, 40], modulus=13)
    with pytest.raises(ValueError):
        gb.ntt([1, 2, 3, 4], modulus=3 * 256 + 2)


def test_live_ntt(live_ntt):
    data, meta = live_ntt
    for key, m in meta.items():
        if "seed" in key:
            continue
        p = int(key.split("/")[0][1:])
        n = int(key.split("/")[1][1:])
        GF = gb.GF(p)
        if n > 1:
            assert int(GF.primitive_root_of_unity(n)) == m["omega"], key
        assert int(GF.primitive_element) == m["primitive_element"]
        for dtype in GF.dtypes:
            if np.dtype(dtype).itemsize < 4 and n > 64:
                continue
            x = GF(data[f"{key}/x"], dtype=dtype)
            X = gb.ntt(x)
            assert type(X) is GF and X.dtype == dtype
            assert np.array_equal(np.asarray(X), data[f"{key}/ntt"]), (key, dtype)
            assert np.array_equal(np.asarray(gb.intt(x)), data[f"{key}/intt"]), (key, dtype)
            assert np.array_equal(np.asarray(gb.intt(x, scaled=False)), data[f"{key}/intt_unscaled"]), (key, dtype)
            assert np.array_equal(np.asarray(np.fft.ifft(np.fft.fft(x))), np.asarray(x))
    GF = gb.GF(998244353)
    assert np.array_equal(np.asarray(gb.ntt(GF(data["pad/x"]), size=128)), data["pad/ntt128"])


@pytest.mark.parametrize("dtype", [np.uint32, np.int64])
def test_live_ntt_2_20(live_ntt, dtype):
    """Headline size N = 2^20 (four-step fast path) against the live reference's digests, host and device."""
    data, meta = live_ntt
    GF = gb.GF(998244353)
    rows = []
    for seed in (3, 4):
        m = meta[f"p998244353/n1048576/seed{seed}"]
        x = GF.Random(2**20, seed=seed, dtype=dtype)
        assert hashlib.sha256(np.asarray(x).astype(np.uint32).tobytes()).hexdigest() == m["x_sha256"]
        X = gb.ntt(x)
        assert hashlib.sha256(np.asarray(X).astype(np.uint32).tobytes()).hexdigest() == m["ntt_sha256"]
        assert np.array_equal(np.asarray(X)[:64], data[f"big/seed{seed}/ntt_head"])
        Xi = gb.intt(x)
        assert hashlib.sha256(np.asarray(Xi).astype(np.uint32).tobytes()).hexdigest() == m["intt_sha256"]
        assert np.array_equal(np.asarray(gb.intt(X)), np.asarray(x))
        rows.append(np.asarray(x))
    # batched (2-D) entry point and device-resident path: rows are independent
    batch = GF(np.stack(rows + rows[::-1] + rows), dtype=dtype)
    Xb = gb.ntt(batch)
    for i, seed in enumerate((3, 4, 4, 3, 3, 4)):
        m = meta[f"p998244353/n1048576/seed{seed}"]
        assert hashlib.sha256(np.asarray(Xb[i]).astype(np.uint32).tobytes()).hexdigest() == m["ntt_sha256"]
    d = batch.to_device()
    Xd = gb.ntt_device(d)
    assert np.array_equal(np.asarray(Xd.cpu()), np.asarray(Xb))
    back = gb.ntt_device(Xd, inverse=True)
    assert np.array_equal(np.asarray(back.cpu()), np.asarray(batch))
    # in-place on the device
    gb.ntt_device(d, out=d)
    assert np.array_equal(np.asarray(d.cpu()), np.asarray(Xb))


def test_ntt_fast_vs_generic_primes():
    """N = 2^20 over other primes (strict p in [2^30, 2^31) and p >= 2^31 generic path) vs the oracle on one row."""
    for p in (2013265921, 469762049, 3221225473):
        GF = gb.GF(p)
        x = GF.Random(2**20, seed=12, dtype=np.int64)
        X = gb.ntt(x)
        omega = int(GF.primitive_root_of_unity(2**20))
        if p < 2**31.5:
            want = oracle.Field.prime(p).fft(np.asarray(x), omega)
            assert np.array_equal(np.asarray(X), want), p
        assert np.array_equal(np.asarray(gb.intt(X)), np.asarray(x)), p
        # linearity: NTT(a + b) = NTT(a) + NTT(b)
        y = GF.Random(2**20, seed=13, dtype=np.int64)
        assert np.array_equal(np.asarray(gb.ntt(x + y)), np.asarray(X + gb.ntt(y))), p


def test_live_linalg(live_linalg):
    d = live_linalg
    groups = sorted({k.rsplit("/", 1)[0] for k in d.files})
    for g in groups:
        parts = g.split("/")
        if parts[0] == "conv":
            GF = make_gf(parts[1])
            z = np.convolve(GF(d[f"{g}/a"]), GF(d[f"{g}/b"]))
            assert np.array_equal(np.asarray(z), d[f"{g}/c"]), g
        elif parts[0] == "matmul":
            GF = make_gf(parts[1])
            z = GF(d[f"{g}/A"]) @ GF(d[f"{g}/B"])
            assert type(z) is GF
            assert np.array_equal(np.asarray(z), d[f"{g}/C"]), g
    GF = gb.GF(2**8)
    A, v = GF.Random((5, 7), seed=1), GF.Random(7, seed=2)
    F = oracle_field("GF(2^8)")
    assert np.array_equal(np.asarray(A @ v), F.matmul(np.asarray(A), np.asarray(v)[:, None])[:, 0])
    assert int(v @ v) == int(F.matmul(np.asarray(v)[None, :], np.asarray(v)[:, None])[0, 0])


def test_ref_vectors_rs_encode(ref_rs):
    GF = gb.GF(2**4, irreducible_poly=0b10011)
    for name in ref_rs["names"]:
        q, n, k, d, alpha, c = (int(v) for v in ref_rs[f"{name}/params"])
        rs = gb.ReedSolomon(n, k, field=GF, alpha=alpha, c=c)
        assert np.array_equal(np.asarray(rs.G), ref_rs[f"{name}/G"]), name
        assert np.array_equal(np.asarray(rs.H), ref_rs[f"{name}/H"]), name
        cw = rs.encode(ref_rs[f"{name}/messages"])
        assert type(cw) is GF and np.array_equal(np.asarray(cw), ref_rs[f"{name}/codewords"]), name
        assert np.array_equal(np.asarray(rs.encode(ref_rs[f"{name}/short_messages"])), ref_rs[f"{name}/short_codewords"])
        assert np.array_equal(np.asarray(rs.encode(ref_rs[f"{name}/messages"][0])), ref_rs[f"{name}/codewords"][0])
        assert not rs.detect(cw).any()


def test_live_rs(live_rs):
    data, meta = live_rs
    for name, m in meta.items():
        GF = gb.GF(2 ** m["m"], irreducible_poly=m["irreducible_poly"], primitive_element=m["primitive_element"])
        rs = gb.ReedSolomon(m["n"], m["k"], field=GF, alpha=m["alpha"], c=m["c"])
        assert np.array_equal(np.asarray(rs.generator_poly), data[f"{name}/generator_poly"]), name
        assert np.array_equal(np.asarray(rs.roots), data[f"{name}/roots"]), name
        assert np.array_equal(np.asarray(rs.G), data[f"{name}/G"]), name
        assert np.array_equal(np.asarray(rs.encode(data[f"{name}/messages"])), data[f"{name}/codewords"]), name
        assert np.array_equal(np.asarray(rs.encode(data[f"{name}/short/messages"])), data[f"{name}/short/codewords"])
        assert np.array_equal(np.asarray(rs.encode(data[f"{name}/messages"], output="parity")),
                              data[f"{name}/codewords"][:, m["k"]:])
        pats = sorted({k[len(name) + 1:].rsplit("/", 1)[0] for k in data.files
                       if k.startswith(name + "/") and k.endswith("/received")})
        for pat in pats:
            R = data[f"{name}/{pat}/received"]
            E = data[f"{name}/{pat}/erasures"]
            dec, nerr = rs.decode(R, erasures=E if E.any() else None, output="codeword", errors=True)
            assert type(dec) is GF and nerr.dtype == np.int64
            assert np.array_equal(nerr, data[f"{name}/{pat}/n_errors"]), (name, pat)
            assert np.array_equal(np.asarray(dec), data[f"{name}/{pat}/decoded"]), (name, pat)
            msg = rs.decode(R, erasures=E if E.any() else None)
            ks = m["k"] - (m["n"] - R.shape[1])
            assert np.array_equal(np.asarray(msg), data[f"{name}/{pat}/decoded"][:, :ks])
        one, ne = rs.decode(data[f"{name}/t_errors/received"][0], errors=True)
        assert isinstance(ne, int) and ne == int(data[f"{name}/t_errors/n_errors"][0])
        assert rs.detect(data[f"{name}/one_error/received"]).all()


def test_rs_255_223_bulk_vs_oracle():
    """BASELINE config 5 shapes at a size the oracle finishes in seconds: patterns A (16 errors), B (32 erasures),
    C (32 errors -> -1), bit-exact vs the oracle; plus the full-size round-trip property on 2e5 codewords."""
    rs = gb.ReedSolomon(255, 223)
    GF = rs.field
    F = oracle.Field.binary(8, 0x11D, alpha=2)
    ors = oracle.ReedSolomon(F, 255, 223)
    N = 3000
    M = GF.Random((N, 223), seed=6)
    C = rs.encode(M)
    assert np.array_equal(np.asarray(C), ors.encode(np.asarray(M)))
    rng = np.random.default_rng(7)

    def corrupt(C, n_err, n_era):
        R = np.asarray(C).copy()
        E = np.zeros(R.shape, dtype=bool)
        for i in range(R.shape[0]):
            pos = rng.choice(255, n_err + n_era, replace=False)
            R[i, pos[:n_err]] ^= rng.integers(1, 256, n_err).astype(np.uint8)
            E[i, pos[n_err:]] = True
            R[i, pos[n_err:]] = rng.integers(0, 256, n_era)
        return R, E

    for n_err, n_era in [(16, 0), (0, 32), (32, 0), (8, 16), (17, 0), (3, 27)]:
        R, E = corrupt(C, n_err, n_era)
        dec, nerr = rs.decode(R, erasures=E if n_era else None, output="codeword", errors=True)
        odec, onerr = ors.decode(R, E if n_era else None, threads=8)
        assert np.array_equal(nerr, onerr), (n_err, n_era)
        assert np.array_equal(np.asarray(dec), odec), (n_err, n_era)
        if 2 * n_err + n_era <= 32:
            assert (nerr == n_err).all() and np.array_equal(np.asarray(dec), np.asarray(C))
        if n_err == 32:
            assert (nerr == -1).all() and np.array_equal(np.asarray(dec), R)
    # round trip at scale through the device-resident path
    N = 200_000
    M = GF.Random((N, 223), seed=66)
    dM = M.to_device()
    dC = rs.encode_device(dM)
    Ch = np.asarray(dC.cpu())
    assert np.array_equal(Ch[:, :223], np.asarray(M))
    R = Ch.copy()
    cols = rng.integers(0, 255, (N, 16))
    R[np.arange(N)[:, None], cols] ^= rng.integers(1, 256, (N, 16)).astype(np.uint8)
    dec, nerr = rs.decode(R, output="codeword", errors=True)
    assert np.array_equal(np.asarray(dec), Ch)
    assert (nerr >= 0).all() and (nerr <= 16).all()


def test_ref_vectors_bch_encode(ref_bch):
    """Sage vectors for the binary BCH(15, k) codes of the reference's test-suite through the CUDA encoder."""
    for name in ref_bch["names"]:
        q, n, k, d, alpha, c = (int(v) for v in ref_bch[f"{name}/params"])
        bch = gb.BCH(n, d=d, alpha=alpha, c=c)
        assert bch.k == k
        assert np.array_equal(np.asarray(bch.G), ref_bch[f"{name}/G"]), name
        cw = bch.encode(ref_bch[f"{name}/messages"])
        assert type(cw) is gb.GF(2) and np.array_equal(np.asarray(cw), ref_bch[f"{name}/codewords"]), name
        if f"{name}/short_messages" in ref_bch.files:
            assert np.array_equal(np.asarray(bch.encode(ref_bch[f"{name}/short_messages"])), ref_bch[f"{name}/short_codewords"])
        if k < n:
            assert not bch.detect(cw).any()


def test_live_bch(live_bch):
    """Binary BCH encode / errors-and-erasures decode vs the live reference, incl. the rows on which the reference
    raises ValueError (miscorrection to non-binary symbols)."""
    data, meta = live_bch
    for name, m in meta.items():
        ext = gb.GF(2 ** m["m"], irreducible_poly=m["irreducible_poly"], primitive_element=m["primitive_element"])
        bch = gb.BCH(m["n"], d=m["d"], extension_field=ext, alpha=m["alpha"], c=m["c"])
        GF2 = bch.field
        assert np.array_equal(np.asarray(bch.G), data[f"{name}/G"]), name
        assert np.array_equal(np.asarray(bch.encode(data[f"{name}/messages"])), data[f"{name}/codewords"]), name
        assert np.array_equal(np.asarray(bch.encode(data[f"{name}/short/messages"])), data[f"{name}/short/codewords"])
        pats = sorted({k[len(name) + 1:].rsplit("/", 1)[0] for k in data.files
                       if k.startswith(name + "/") and k.endswith("/received")})
        for pat in pats:
            R, E = data[f"{name}/{pat}/received"], data[f"{name}/{pat}/erasures"]
            raised = data[f"{name}/{pat}/raised"]
            ok = ~raised
            if ok.any():
                dec, nerr = bch.decode(R[ok], erasures=E[ok] if E.any() else None, output="codeword", errors=True)
                assert type(dec) is GF2 and nerr.dtype == np.int64
                assert np.array_equal(nerr, data[f"{name}/{pat}/n_errors"][ok]), (name, pat)
                assert np.array_equal(np.asarray(dec), data[f"{name}/{pat}/decoded"][ok]), (name, pat)
            for i in np.flatnonzero(raised)[:3]:
                with pytest.raises(ValueError, match="GF\\(2\\) arrays must have elements"):
                    bch.decode(R[i], erasures=E[i] if E.any() else None)


def test_bch_255_131_bulk_vs_oracle():
    """BCH(255,131,t=18): 20000 codewords with 18 bit errors each vs the oracle, and a 2e5-codeword round trip."""
    bch = gb.BCH(255, 131)
    assert bch.d == 37
    F = oracle.Field.binary(8, 0x11D, alpha=2)
    ob = oracle.BCH(F, 255, 37)
    N = 20000
    M = bch.field.Random((N, 131), seed=1)
    C = bch.encode(M)
    assert np.array_equal(np.asarray(C), ob.encode(np.asarray(M), threads=8))
    rng = np.random.default_rng(2)
    R = np.asarray(C).copy()
    cols = np.argsort(rng.random((N, 255)), axis=1)[:, :18]
    R[np.arange(N)[:, None], cols] ^= 1
    dec, nerr = bch.decode(R, output="codeword", errors=True)
    odec, onerr = ob.decode(R, threads=8)
    assert np.array_equal(nerr, onerr) and (nerr == 18).all()
    assert np.array_equal(np.asarray(dec), odec) and np.array_equal(np.asarray(dec), np.asarray(C))
    N = 200_000
    M = bch.field.Random((N, 131), seed=3)
    C = np.asarray(bch.encode(M))
    R = C.copy()
    cols = np.argsort(rng.random((N, 255)), axis=1)[:, :18]
    R[np.arange(N)[:, None], cols] ^= 1
    dec, nerr = bch.decode(R, output="codeword", errors=True)
    assert np.array_equal(np.asarray(dec), C) and (nerr == 18).all()
    assert np.array_equal(np.asarray(bch.decode(R)), np.asarray(M))


def test_poly_evaluate(live_linalg):
    """Horner evaluation (evaluate_elementwise_jit behind Poly.__call__) vs the live
    reference's Poly(...)(x) and vs the oracle on larger cases over several fields."""
    d = live_linalg
    GF = gb.GF(2**8)
    n_cases = len({k.split("/")[1] for k in d.files if k.startswith("eval/")})
    assert n_cases >= 10
    for i in range(n_cases):
        y = gb.poly_evaluate(GF(d[f"eval/{i}/coeffs"]), GF(d[f"eval/{i}/x"]))
        assert type(y) is GF and np.array_equal(np.asarray(y), d[f"eval/{i}/y"]), i
    for name in ("GF(2147483647)", "GF(998244353)", "GF(2^8)", "GF(2^32)", "GF(7)"):
        G = make_gf(name)
        F = oracle_field(name)
        for nc, n in [(1, 5), (2, 1000), (33, 4097), (1500, 20000)]:
            c = G.Random(nc, seed=nc)
            x = G.Random(n, seed=n + 1)
            y = gb.poly_evaluate(c, x)
            assert type(y) is G and y.dtype == x.dtype and y.shape == x.shape
            assert np.array_equal(np.asarray(y).astype(np.int64), F.poly_evaluate(np.asarray(c), np.asarray(x))), (name, nc, n)
    # multipoint evaluation at the N-th roots of unity is the NTT (tests/fields/test_fft.py:33-45)
    G = gb.GF(998244353)
    x = G.Random(256, seed=5)
    w = int(G.primitive_root_of_unity(256))
    pts = G([pow(w, k, G.order) for k in range(256)])
    assert np.array_equal(np.asarray(gb.poly_evaluate(G(np.asarray(x)[::-1].copy()), pts)), np.asarray(gb.ntt(x)))


@pytest.mark.parametrize("log_n", [15, 16, 17, 18, 19, 20])
def test_ntt_four_step_sizes(log_n):
    """Every size of the four-step fast path (N = n1 x 1024, n1 = 32..1024) vs the oracle, batched, forward and
    inverse, 4- and 8-byte storage, lazy (p < 2^30) and strict (2^30 <= p < 2^31) arithmetic."""
    n = 1 << log_n
    for p in (998244353, 2013265921):
        GF = gb.GF(p)
        F = oracle.Field.prime(p)
        omega = int(GF.primitive_root_of_unity(n))
        for dtype in (np.uint32, np.int64):
            x = GF.Random((3, n), seed=log_n, dtype=dtype)
            X = gb.ntt(x)
            assert type(X) is GF and X.dtype == dtype
            assert np.array_equal(np.asarray(X).astype(np.int64), F.fft(np.asarray(x), omega, threads=3)), (log_n, p, dtype)
            assert np.array_equal(np.asarray(gb.intt(X)), np.asarray(x)), (log_n, p, dtype)
        d = x.to_device()
        assert np.array_equal(np.asarray(gb.ntt_device(gb.ntt_device(d), inverse=True).cpu()), np.asarray(x))


@pytest.mark.parametrize("log_n", list(range(0, 15)))
def test_ntt_small_sizes(log_n):
    """N = 1 .. 2^14: generic Stockham path (N < 32), single-kernel path (32 .. 1024) and short-column four-step
    (2^11 .. 2^14) vs the oracle, with batch sizes that are not multiples of the 16-transform tile."""
    n = 1 << log_n
    for p in (998244353, 2013265921, 7340033):
        if (p - 1) % n:
            continue
        GF = gb.GF(p)
        F = oracle.Field.prime(p)
        omega = int(GF.primitive_root_of_unity(n)) if n > 1 else 1
        for rows in (1, 5, 16, 37):
            x = GF.Random((rows, n), seed=100 + log_n + rows)
            X = gb.ntt(x)
            assert np.array_equal(np.asarray(X).astype(np.int64), F.fft(np.asarray(x), omega)), (log_n, p, rows)
            assert np.array_equal(np.asarray(gb.intt(X)), np.asarray(x)), (log_n, p, rows)
        x64 = GF.Random((3, n), seed=7, dtype=np.int64)
        assert np.array_equal(np.asarray(gb.ntt(x64)).astype(np.int64), F.fft(np.asarray(x64), omega))
        assert np.array_equal(np.asarray(gb.intt(x64, scaled=False)).astype(np.int64),
                              F.fft(np.asarray(x64), pow(omega, -1, p)))


@pytest.mark.parametrize(["n", "k"], [(255, 155), (255, 101), (255, 5), (63, 3)])
def test_rs_low_rate_codes_vs_oracle(n, k):
    """Low-rate codes: d - 1 = 100 .. 250 syndromes exercise the multi-register Berlekamp-Massey (4 and 8 coefficients
    per lane) and long erasure locators; errors, erasures, mixtures and failures vs the oracle."""
    m = 8 if n == 255 else 6
    poly = {8: 0x11D, 6: 0x43}[m]
    rs = gb.ReedSolomon(n, k)
    assert int(rs.field.irreducible_poly) == poly
    F = oracle.Field.binary(m, poly, alpha=2)
    ors = oracle.ReedSolomon(F, n, k)
    d1 = n - k
    rng = np.random.default_rng(n * 1000 + k)
    N = 300
    M = rs.field.Random((N, k), seed=k)
    C = np.asarray(rs.encode(M))
    assert np.array_equal(C, ors.encode(np.asarray(M)))
    for n_err, n_era in [(d1 // 2, 0), (0, d1), (d1 // 4, d1 - 2 * (d1 // 4)), (d1 // 2 + 1, 0), (1, d1), (3, 5)]:
        R = C.copy()
        E = np.zeros(R.shape, dtype=bool)
        for i in range(N):
            pos = rng.choice(n, min(n, n_err + n_era), replace=False)
            R[i, pos[:n_err]] ^= rng.integers(1, 1 << m, len(pos[:n_err])).astype(np.uint8)
            E[i, pos[n_err:]] = True
            R[i, pos[n_err:]] = rng.integers(0, 1 << m, len(pos[n_err:]))
        dec, nerr = rs.decode(R, erasures=E if n_era else None, output="codeword", errors=True)
        odec, onerr = ors.decode(R, E if n_era else None, threads=8)
        assert np.array_equal(nerr, onerr), (n, k, n_err, n_era)
        assert np.array_equal(np.asarray(dec), odec), (n, k, n_err, n_era)


@pytest.mark.parametrize("name", ["GF(2^8)", "GF(31)", "GF(2147483647)", "GF(2^32)", "GF(3191)"])
def test_ufunc_methods_reduce_accumulate_reduceat_at(name):
    """np.<ufunc>.reduce / accumulate / reduceat / at (the reference forwards them to its DUFunc, _ufunc.py:349,403;
    tests/fields/test_numpy_ufuncs.py): serial folds with the oracle's field operations as the expected values."""
    GF = make_gf(name)
    F = oracle_field(name)
    rng = np.random.default_rng(5)
    order = GF.order
    ops = {np.add: F.add, np.subtract: F.sub, np.multiply: F.mul, np.true_divide: F.div}

    def serial(fn, rows):  # running left fold along the last axis
        rows = np.asarray(rows, dtype=np.int64)
        acc = rows[..., 0].copy()
        out = [acc.copy()]
        for j in range(1, rows.shape[-1]):
            acc = np.asarray(fn(acc.reshape(-1), rows[..., j].reshape(-1))).reshape(acc.shape)
            out.append(acc.copy())
        return np.stack(out, axis=-1)

    for shape in [(7,), (1,), (3, 41), (5, 4, 6), (70000,), (2, 40000)]:
        x = rng.integers(1, order, shape)
        gx = GF(x)
        for ufunc, fn in ops.items():
            for axis in range(len(shape)):
                want = np.moveaxis(serial(fn, np.moveaxis(x, axis, -1)), -1, axis)
                got = ufunc.accumulate(gx, axis=axis)
                assert type(got) is GF and got.dtype == gx.dtype
                assert eq(got, want), (name, ufunc.__name__, shape, axis, "accumulate")
                red = ufunc.reduce(gx, axis=axis)
                assert eq(red, np.take(want, -1, axis=axis)), (name, ufunc.__name__, shape, axis, "reduce")
    x = rng.integers(1, order, 50)
    gx = GF(x)
    idx = np.array([0, 7, 7, 3, 20, 49, 10])
    for ufunc, fn in ops.items():
        want = []
        for i, s in enumerate(idx):
            e = idx[i + 1] if i + 1 < len(idx) else len(x)
            seg = x[s:e] if e > s else x[s:s + 1]
            want.append(serial(fn, seg[None, :])[0, -1])
        got = ufunc.reduceat(gx, idx)
        assert type(got) is GF and eq(got, np.array(want)), (name, ufunc.__name__, "reduceat")
    X2 = rng.integers(1, order, (4, 30))
    got = np.add.reduceat(GF(X2), [0, 5, 5, 29], axis=1)
    want = np.stack([serial(F.add, X2[:, 0:5])[:, -1], X2[:, 5], serial(F.add, X2[:, 5:29])[:, -1], X2[:, 29]], axis=1)
    assert eq(got, want)
    # at: repeated indices are applied once per occurrence, in order
    for ufunc, fn in ops.items():
        a = x.copy()
        ga = GF(x)
        ii = np.array([3, 3, 0, 49, 3, 7])
        b = rng.integers(1, order, ii.size)
        assert ufunc.at(ga, ii, GF(b)) is None
        for k, i in enumerate(ii):
            a[i] = fn(np.array([a[i]]), np.array([b[k]]))[0]
        assert eq(ga, a), (name, ufunc.__name__, "at")
    ga = GF(x)
    np.negative.at(ga, [1, 1, 2])
    want = x.copy()
    want[2] = F.neg(np.array([x[2]]))[0]
    assert eq(ga, want)
    ga = GF(x)
    np.power.at(ga, [0, 0, 5], np.array([3, 2, -1]))
    want = x.copy()
    want[0] = F.power(np.array([x[0]]), np.array([6]))[0]
    want[5] = F.inv(np.array([x[5]]))[0]
    assert eq(ga, want)
    with pytest.raises(ValueError):
        np.negative.reduce(gx)
    with pytest.raises(ValueError):
        np.power.accumulate(gx)


@pytest.mark.gpu
@pytest.mark.parametrize("ns,offset,N", [(255, 1, 97), (255, 3, 64), (200, 2, 45), (64, 1, 33), (33, 0, 32)])
def test_rs_decode_unaligned_buffers_vs_oracle(ns, offset, N):
    """The bit-sliced decoder reads each row in aligned words: device buffers that start at any byte offset, shortened codes
    (rows of ns < n symbols), a partial last group and erasures must give the oracle's rows bit for bit."""
    import torch

    from galois_b200 import _lib
    from galois_b200._field import DeviceArray

    rs = gb.ReedSolomon(255, 223)
    F = oracle.Field.binary(8, 0x11D, alpha=2)
    ors = oracle.ReedSolomon(F, 255, 223)
    ks = ns - 32
    rng = np.random.default_rng(1000 + ns + offset)
    M = rng.integers(0, 256, (N, ks), dtype=np.uint8)
    C = ors.encode(M)
    R = C.copy()
    E = np.zeros(R.shape, dtype=bool)
    for i in range(N):
        n_era = int(rng.integers(0, 9)) if i % 3 else 0
        n_err = int(rng.integers(0, (32 - n_era) // 2 + 2))  # some rows beyond the correction radius
        pos = rng.choice(ns, min(ns, n_err + n_era), replace=False)
        R[i, pos[:n_err]] ^= rng.integers(1, 256, len(pos[:n_err])).astype(np.uint8)
        E[i, pos[n_err:]] = True
    for use_era in (False, True):
        odec, onerr = ors.decode(R, E if use_era else None, threads=4)
        buf_r = torch.zeros(N * ns + 8, dtype=torch.uint8, device="cuda")
        buf_e = torch.zeros(N * ns + 8, dtype=torch.uint8, device="cuda")
        buf_o = torch.zeros(N * ns + 8, dtype=torch.uint8, device="cuda")
        buf_r[offset:offset + N * ns] = torch.from_numpy(R.reshape(-1)).cuda()
        buf_e[offset + 1:offset + 1 + N * ns] = torch.from_numpy(E.reshape(-1).astype(np.uint8)).cuda()
        nerr = torch.empty(N, dtype=torch.int64, device="cuda")
        st = _lib.lib().gfb_rs_decode(rs._code(), buf_r.data_ptr() + offset, buf_e.data_ptr() + offset + 1 if use_era else None, ns,
                                      buf_o.data_ptr() + offset, nerr.data_ptr(), N, DeviceArray._stream())
        _lib.check(st, "gfb_rs_decode")
        torch.cuda.synchronize()
        got = buf_o[offset:offset + N * ns].cpu().numpy().reshape(N, ns)
        assert np.array_equal(nerr.cpu().numpy(), onerr), (ns, offset, use_era)
        assert np.array_equal(got, odec), (ns, offset, use_era)
        assert int(buf_o[:offset].sum()) == 0 and int(buf_o[offset + N * ns:].sum()) == 0  # nothing written outside the rows
