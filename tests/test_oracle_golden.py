"""
Pins oracle/ (the CPU restatement) against the reference:
  * the reference's own Sage/SymPy vectors (ref_*.npz; /root/reference/tests/fields/data, tests/codes/data),
  * outputs of the live reference on seeded inputs (live_*.npz; tests/golden/make_golden.py).
No GPU needed.
"""

import numpy as np
import pytest

import oracle
from helpers import oracle_field, parse_field

REF_FIELDS = ["GF(2)", "GF(2^2)", "GF(2^3)", "GF(2^8)", "GF(2^32)", "GF(2^100)", "GF(5)", "GF(7)", "GF(31)", "GF(3191)",
              "GF(2147483647)"]
NARROW = [f for f in REF_FIELDS if f != "GF(2^100)"]


def _modes(name):
    spec = parse_field(name)
    if spec["p"] == 2 and spec["m"] <= 16:
        return [True, False]  # jit-lookup and jit-calculate must agree (tests/fields/conftest.py:17-72)
    return [None]


@pytest.mark.parametrize("name", NARROW)
def test_ref_vectors_arithmetic(ref_fields, name):
    """tests/fields/test_arithmetic.py vs Sage: add, subtract, multiply, divide, inverses, power."""
    data, _ = ref_fields
    for lookup in _modes(name):
        F = oracle_field(name, lookup)
        for op, fn in [("add", F.add), ("subtract", F.sub), ("multiply", F.mul), ("divide", F.div)]:
            X, Y = np.meshgrid(data[f"{name}/{op}/X"], data[f"{name}/{op}/Y"], indexing="ij")
            assert np.array_equal(fn(X, Y), data[f"{name}/{op}/Z"]), (name, op, lookup)
        assert np.array_equal(F.neg(data[f"{name}/additive_inverse/X"]), data[f"{name}/additive_inverse/Z"])
        assert np.array_equal(F.inv(data[f"{name}/multiplicative_inverse/X"]), data[f"{name}/multiplicative_inverse/Z"])
        X, Y = np.meshgrid(data[f"{name}/power/X"], data[f"{name}/power/Y"], indexing="ij")
        assert np.array_equal(F.power(X, Y), data[f"{name}/power/Z"]), (name, "power", lookup)


def test_ref_vectors_gf2_100(ref_fields):
    """The reference's widest pinned binary field (python-calculate, object dtype)."""
    data, _ = ref_fields
    name = "GF(2^100)"
    F = oracle_field(name)
    for op, code in [("add", oracle.ADD), ("subtract", oracle.SUB), ("multiply", oracle.MUL), ("divide", oracle.DIV)]:
        X, Y, Z = data[f"{name}/{op}/X"], data[f"{name}/{op}/Y"], data[f"{name}/{op}/Z"]
        nx, ny = X.shape[0], Y.shape[0]
        xx = np.repeat(X, ny, axis=0)
        yy = np.tile(Y, (nx, 1))
        assert np.array_equal(F.binary_op(code, xx, yy).reshape(nx, ny, 2), Z), op
    X, Z = data[f"{name}/multiplicative_inverse/X"], data[f"{name}/multiplicative_inverse/Z"]
    assert np.array_equal(F.inv(X), Z)
    X, Y, Z = data[f"{name}/power/X"], data[f"{name}/power/Y"], data[f"{name}/power/Z"]
    xx = np.repeat(X, Y.shape[0], axis=0)
    ee = np.tile(Y, X.shape[0])
    assert np.array_equal(F.power(xx, ee).reshape(Z.shape), Z)


@pytest.mark.parametrize("name", NARROW)
def test_ref_vectors_convolve_matmul(ref_fields, name):
    """tests/fields/test_advanced_arithmetic.py:10-17 and tests/fields/test_linalg.py:275-280 vs Sage."""
    data, _ = ref_fields
    F = oracle_field(name)
    for i in range(int(data[f"{name}/convolve/count"])):
        z = F.convolve(data[f"{name}/convolve/X{i}"], data[f"{name}/convolve/Y{i}"])
        assert np.array_equal(z, data[f"{name}/convolve/Z{i}"])
    for i in range(int(data[f"{name}/matrix_multiply/count"])):
        z = F.matmul(data[f"{name}/matrix_multiply/X{i}"], data[f"{name}/matrix_multiply/Y{i}"])
        assert np.array_equal(z, data[f"{name}/matrix_multiply/Z{i}"])


# SymPy LUTs, /root/reference/tests/fields/test_ntt.py:13-18
NTT_LUTS = [
    ([1, 2, 3, 4], 5, [0, 4, 3, 2], 2),
    ([1, 2, 3, 4], 13, [10, 8, 11, 1], 8),
    ([1, 2, 3, 4], 17, [10, 6, 15, 7], 13),
    ([1, 2, 3, 4], 3 * 256 + 1, [10, 643, 767, 122], 62),
]


@pytest.mark.parametrize(["x", "p", "X", "omega"], NTT_LUTS)
def test_ref_vectors_ntt_sympy(x, p, X, omega):
    F = oracle.Field.prime(p)
    assert list(F.fft(x, omega)) == X


def test_ref_vectors_rs_encode(ref_rs):
    """Sage RS(15,k)/GF(2^4) vectors, /root/reference/tests/codes/conftest.py:115-171."""
    F = oracle.Field.binary(4, 0b10011, alpha=2)
    for name in ref_rs["names"]:
        q, n, k, d, alpha, c = (int(v) for v in ref_rs[f"{name}/params"])
        assert q == 16
        rs = oracle.ReedSolomon(F, n, k, c=c, alpha=alpha)
        G = ref_rs[f"{name}/G"]
        assert np.array_equal(rs.P, G[:, k:]), name
        assert np.array_equal(rs.encode(ref_rs[f"{name}/messages"]), ref_rs[f"{name}/codewords"]), name
        assert np.array_equal(rs.encode(ref_rs[f"{name}/short_messages"]), ref_rs[f"{name}/short_codewords"]), name


def test_live_elementwise(live_elementwise):
    data, _ = live_elementwise
    keys = sorted({k.rsplit("/", 1)[0] for k in data.files})
    assert keys
    for key in keys:
        name = key.split("/")[0]
        for lookup in _modes(name):
            F = oracle_field(name, lookup)
            g = lambda s: data[f"{key}/{s}"].astype(np.int64)  # noqa: E731
            x, y, ynz = g("x"), g("y"), g("ynz")
            assert np.array_equal(F.add(x, y), g("add")), key
            assert np.array_equal(F.sub(x, y), g("sub")), key
            assert np.array_equal(F.mul(x, y), g("mul")), key
            assert np.array_equal(F.div(x, ynz), g("div")), key
            assert np.array_equal(F.neg(x), g("neg")), key
            assert np.array_equal(F.inv(ynz), g("inv")), key
            assert np.array_equal(F.square(x), g("square")), key
            assert np.array_equal(F.power(g("xe"), g("e")), g("pow")), key
            assert np.array_equal(F.power(x, 3), g("pow3")), key
            assert np.array_equal(F.power(ynz, -1), g("powm1")), key
            order = F.order
            assert np.array_equal(F.mul(x, 7 % order), g("mul_scalar7")), key
            assert np.array_equal(F.sub(5 % order, x), g("scalar5_sub")), key
            assert np.array_equal(F.div(3 % order, ynz), g("scalar3_div")), key
            with pytest.raises(ZeroDivisionError):
                F.inv(np.array([1, 0]))


def test_live_gf2_128(live_gf128):
    d = live_gf128
    F = oracle.Field128(128, (1 << 128) | 0x87)
    assert np.array_equal(F.mul(d["a"], d["b"]), d["mul"])
    assert np.array_equal(F.add(d["a"], d["b"]), d["add"])
    assert np.array_equal(F.unary_op(oracle.SQUARE, d["a"]), d["square"])
    assert np.array_equal(F.inv(d["inv_in"]), d["inv"])
    n = d["inv_in"].shape[0]
    assert np.array_equal(F.div(d["a"][:n], d["inv_in"]), d["div"])
    assert np.array_equal(F.power(d["pow_in"], d["pow_e"]), d["pow"])
    poly100 = int(d["gf100_poly"][0, 0]) | (int(d["gf100_poly"][0, 1]) << 64)
    F100 = oracle.Field128(100, poly100)
    assert np.array_equal(F100.mul(d["gf100_a"], d["gf100_b"]), d["gf100_mul"])
    assert np.array_equal(F100.inv(d["gf100_inv_in"]), d["gf100_inv"])


def test_live_ntt(live_ntt):
    data, meta = live_ntt
    for key, m in meta.items():
        if "seed" in key:
            continue
        p = int(key.split("/")[0][1:])
        n = int(key.split("/")[1][1:])
        F = oracle.Field.prime(p)
        omega = m["omega"]
        x = data[f"{key}/x"]
        assert np.array_equal(F.fft(x, omega), data[f"{key}/ntt"]), key
        omega_inv = int(F.inv(np.array([omega]))[0])
        n_inv = int(F.inv(np.array([n % p]))[0])
        assert np.array_equal(F.fft(x, omega_inv), data[f"{key}/intt_unscaled"]), key
        assert np.array_equal(F.fft(x, omega_inv, scale=n_inv), data[f"{key}/intt"]), key


def test_live_ntt_2_20(live_ntt):
    """The headline transform size, pinned by SHA-256 of the live reference's output."""
    import hashlib

    data, meta = live_ntt
    F = oracle.Field.prime(998244353)
    m = meta["p998244353/n1048576/seed3"]
    x = np.random.default_rng(3).integers(0, 998244353, 2**20, dtype=np.uint32)
    assert hashlib.sha256(x.tobytes()).hexdigest() == m["x_sha256"]
    X = F.fft(x.astype(np.int64), m["omega"]).astype(np.uint32)
    assert hashlib.sha256(X.tobytes()).hexdigest() == m["ntt_sha256"]
    assert np.array_equal(X[:: 2**12].astype(np.int64), data["big/seed3/ntt_stride"])


def test_live_linalg(live_linalg):
    d = live_linalg
    groups = sorted({k.rsplit("/", 1)[0] for k in d.files})
    for g in groups:
        kind, name = g.split("/")[0], g.split("/")[1] if "/" in g else None
        if kind == "conv":
            F = oracle_field(name)
            assert np.array_equal(F.convolve(d[f"{g}/a"], d[f"{g}/b"]), d[f"{g}/c"]), g
        elif kind == "matmul":
            F = oracle_field(name)
            assert np.array_equal(F.matmul(d[f"{g}/A"], d[f"{g}/B"]), d[f"{g}/C"]), g
        elif kind == "bm":
            F = oracle_field("GF(2^8)")
            # the fixture holds galois.berlekamp_massey()'s Poly = the jit kernel's C(x) reversed (_lfsr.py:1560-1619)
            assert np.array_equal(F.berlekamp_massey(d[f"{g}/s"])[::-1], d[f"{g}/c"]), g
        elif kind == "eval":
            F = oracle_field("GF(2^8)")
            assert np.array_equal(F.poly_evaluate(d[f"{g}/coeffs"], d[f"{g}/x"]), d[f"{g}/y"]), g


def test_live_rs(live_rs):
    data, meta = live_rs
    for name, m in meta.items():
        F = oracle.Field.binary(m["m"], m["irreducible_poly"], alpha=m["primitive_element"])
        rs = oracle.ReedSolomon(F, m["n"], m["k"], c=m["c"], alpha=m["alpha"])
        assert np.array_equal(rs.generator_poly, data[f"{name}/generator_poly"]), name
        assert np.array_equal(rs.roots, data[f"{name}/roots"]), name
        assert np.array_equal(rs.P, data[f"{name}/G"][:, m["k"]:]), name
        assert np.array_equal(rs.encode(data[f"{name}/messages"]), data[f"{name}/codewords"]), name
        assert np.array_equal(rs.encode(data[f"{name}/short/messages"]), data[f"{name}/short/codewords"]), name
        pats = sorted({k[len(name) + 1:].rsplit("/", 1)[0] for k in data.files if k.startswith(name + "/") and k.endswith("/received")})
        assert pats
        for pat in pats:
            R = data[f"{name}/{pat}/received"]
            E = data[f"{name}/{pat}/erasures"]
            dec, nerr = rs.decode(R, E if E.any() else None)
            assert np.array_equal(nerr, data[f"{name}/{pat}/n_errors"]), (name, pat)
            assert np.array_equal(dec, data[f"{name}/{pat}/decoded"]), (name, pat)


def test_ref_vectors_bch_encode(ref_bch):
    """Sage vectors for binary BCH(15, k) codes, /root/reference/tests/codes/data/bch/n15_*_sys.pkl (q = 2)."""
    F = oracle.Field.binary(4, 0b10011, alpha=2)
    assert len(ref_bch["names"]) >= 10
    for name in ref_bch["names"]:
        q, n, k, d, alpha, c = (int(v) for v in ref_bch[f"{name}/params"])
        bch = oracle.BCH(F, n, d, c=c, alpha=alpha)
        assert bch.k == k, name
        assert np.array_equal(bch.P, ref_bch[f"{name}/G"][:, k:]), name
        assert np.array_equal(bch.encode(ref_bch[f"{name}/messages"]), ref_bch[f"{name}/codewords"]), name
        if f"{name}/short_messages" in ref_bch.files:
            assert np.array_equal(bch.encode(ref_bch[f"{name}/short_messages"]), ref_bch[f"{name}/short_codewords"]), name


def test_live_bch(live_bch):
    data, meta = live_bch
    for name, m in meta.items():
        F = oracle.Field.binary(m["m"], m["irreducible_poly"], alpha=m["primitive_element"])
        bch = oracle.BCH(F, m["n"], m["d"], c=m["c"], alpha=m["alpha"])
        assert bch.k == m["k"], name
        assert np.array_equal(bch.generator_poly, data[f"{name}/generator_poly"]), name
        assert np.array_equal(bch.roots, data[f"{name}/roots"]), name
        assert np.array_equal(bch.P, data[f"{name}/G"][:, m["k"]:]), name
        assert np.array_equal(bch.encode(data[f"{name}/messages"]), data[f"{name}/codewords"]), name
        assert np.array_equal(bch.encode(data[f"{name}/short/messages"]), data[f"{name}/short/codewords"]), name
        pats = sorted({k[len(name) + 1:].rsplit("/", 1)[0] for k in data.files if k.startswith(name + "/") and k.endswith("/received")})
        assert len(pats) >= 8
        for pat in pats:
            R, E = data[f"{name}/{pat}/received"], data[f"{name}/{pat}/erasures"]
            raised = data[f"{name}/{pat}/raised"]
            dec, nerr = bch.decode(R, E if E.any() else None)
            ok = ~raised
            assert np.array_equal(nerr[ok], data[f"{name}/{pat}/n_errors"][ok]), (name, pat)
            assert np.array_equal(dec[ok], data[f"{name}/{pat}/decoded"][ok]), (name, pat)
            # rows on which the reference raised ValueError are exactly the rows decoded to non-binary symbols
            assert np.array_equal((dec >= 2).any(axis=1), raised), (name, pat)
