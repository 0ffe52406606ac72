"""
GPU parity tests (``-m gpu``) for the SURVEY.md §8(f) rows — Poly.roots / Poly.__call__, the mixed-radix DFT and
dense linear algebra — through the NumPy-facing mirror and hence the C ABI, against
  * the reference's Sage vectors (ref_roots.npz, ref_linalg.npz),
  * outputs of the live reference (live_next.npz),
  * the oracle on seeded inputs, and size-independent properties at sizes the oracle would take too long for.
Bit-exact everywhere.
"""

import numpy as np
import pytest

import galois_b200 as gb
import oracle
from helpers import oracle_field, parse_field

pytestmark = pytest.mark.gpu

FIELDS = ["GF(2)", "GF(2^2)", "GF(2^3)", "GF(2^8)", "GF(2^32)", "GF(5)", "GF(7)", "GF(31)", "GF(3191)", "GF(2147483647)"]


def make_gf(name):
    spec = parse_field(name)
    if spec["p"] != 2:
        return gb.GF(spec["p"])
    return gb.GF(2 ** spec["m"], irreducible_poly=spec["poly"], primitive_element=spec["alpha"])


def live_gf(meta):
    q = meta["order"]
    if q > 2 and q & (q - 1) == 0:
        return gb.GF(q, irreducible_poly=meta["irreducible_poly"], primitive_element=meta["primitive_element"])
    return gb.GF(q)


def live_oracle(meta):
    q = meta["order"]
    if q > 2 and q & (q - 1) == 0:
        m = q.bit_length() - 1
        return oracle.Field.binary(m, meta["irreducible_poly"], alpha=meta["primitive_element"] if m <= 16 else None)
    return oracle.Field.prime(q)


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and np.array_equal(a.astype(np.int64), b.astype(np.int64))


# ------------------------------------------------------------------------------------------------ roots / evaluation
@pytest.mark.parametrize("name", FIELDS)
def test_ref_vectors_roots(ref_roots, name):
    """tests/polys/test_arithmetic_methods.py:28-43 vs Sage, including GF(2^32) and GF(2^31 - 1) (4e9 / 2e9 points)."""
    d = ref_roots
    GF = make_gf(name)
    for i in range(int(d[f"{name}/count"])):
        coeffs = d[f"{name}/X{i}"]
        if not coeffs.any():
            continue
        for dtype in (GF.dtypes[0], GF.dtypes[-1]):
            p = gb.Poly(GF(coeffs, dtype=dtype))
            r = p.roots()
            assert type(r) is GF and r.dtype == dtype
            assert eq(r, d[f"{name}/R{i}"]), (name, i)
            if r.size:
                assert not np.asarray(p(r)).any()  # f(root) == 0 through the evaluation kernel


def test_live_roots(live_next):
    data, meta = live_next
    for name in data["roots/names"]:
        fname = str(name).split("/")[0]
        GF = live_gf(meta["roots"][fname])
        degrees, coeffs = data[f"roots/{name}/degrees"], data[f"roots/{name}/coeffs"]
        r = gb.poly_roots(degrees, GF(coeffs))
        assert type(r) is GF and eq(r, data[f"roots/{name}/roots"]), name
        # the Poly holder gives the same answer, whatever order the terms are given in
        perm = np.random.default_rng(1).permutation(degrees.size)
        assert eq(gb.Poly.Degrees(degrees[perm], coeffs[perm], field=GF).roots(), data[f"roots/{name}/roots"])


def test_roots_vs_oracle_and_properties():
    rng = np.random.default_rng(99)
    # every element of GF(p) is a root of x^p - x; none of x^(p-1) - 1 is 0
    GF = gb.GF(65537)
    r = gb.Poly.Degrees([65537, 1], [1, -1], field=GF).roots()
    assert eq(r, np.arange(65537))
    r = gb.Poly.Degrees([65536, 0], [1, -1], field=GF).roots()
    assert eq(r, np.arange(1, 65537))
    # product of distinct linear factors over a 31-bit prime: the roots come back sorted
    GF = gb.GF(998244353)
    F = oracle.Field.prime(998244353)
    roots = np.unique(rng.integers(0, 998244353, 40))
    poly = np.array([1], dtype=np.int64)
    for z in roots:
        poly = F.convolve(poly, np.array([1, (998244353 - z) % 998244353]))
    got = gb.Poly(GF(poly)).roots()
    assert eq(got, roots)
    # a long dense polynomial (several coefficient chunks) vs the oracle over GF(2^12)
    GF = gb.GF(2**12)
    F12 = oracle.Field.binary(12, int(GF.irreducible_poly), alpha=int(GF.primitive_element))
    c = rng.integers(0, 4096, 1500)
    c[0] = 1
    nz = np.nonzero(c)[0]
    want = F12.poly_roots((c.size - 1 - nz).astype(np.int64), c[nz], primitive_element=int(GF.primitive_element))
    assert eq(gb.Poly(GF(c)).roots(), want)
    with pytest.raises(ValueError):
        gb.poly_roots(np.array([], dtype=np.int64), GF([]))


def test_poly_call_matches_oracle():
    GF = gb.GF(3191)
    F = oracle.Field.prime(3191)
    c = GF.Random(33, seed=1)
    x = GF.Random((7, 11), seed=2)
    p = gb.Poly(c)
    y = p(x)
    assert type(y) is GF and y.shape == x.shape
    assert eq(y, F.poly_evaluate(np.asarray(c), np.asarray(x).reshape(-1)).reshape(x.shape))
    assert int(p(5)) == int(F.poly_evaluate(np.asarray(c), np.array([5]))[0])


# ------------------------------------------------------------------------------------------------ mixed-radix DFT
def test_live_fft_mixed_radix(live_next):
    data, meta = live_next
    for fname, m in meta["fft"].items():
        GF = live_gf(m)
        for n, omega in zip(m["sizes"], m["omegas"]):
            if n > 1:
                assert int(GF.primitive_root_of_unity(n)) == omega
            x = GF(data[f"fft/{fname}/{n}/x"])
            X = np.fft.fft(x)
            assert type(X) is GF and X.dtype == x.dtype
            assert eq(X, data[f"fft/{fname}/{n}/X"]), (fname, n)
            assert eq(np.fft.ifft(x), data[f"fft/{fname}/{n}/ix"]), (fname, n)
            assert eq(np.fft.ifft(X), x)


def test_fft_mixed_batched_vs_oracle():
    """2-D batches, zero-padding / truncation via n=, 8-byte elements and the device-resident entry point."""
    for GF, F, sizes in [
        (gb.GF(998244353), oracle.Field.prime(998244353), [17, 7 * 17 * 4, 119, 17 * 512]),
        (gb.GF(2**31 - 1), oracle.Field.prime(2**31 - 1), [331, 9 * 7 * 11, 331 * 6]),
        (gb.GF(2**16), None, [257, 255, 257 * 5]),
    ]:
        if F is None:
            F = oracle.Field.binary(16, int(GF.irreducible_poly), alpha=int(GF.primitive_element))
        for n in sizes:
            omega = int(GF.primitive_root_of_unity(n))
            x = GF.Random((5, n), seed=n)
            want = F.fft(np.asarray(x), omega)
            assert eq(np.fft.fft(x), want), (GF.name, n)
            x64 = x.astype(GF.dtypes[-1])
            assert eq(np.fft.fft(x64), want)
            # n= pads with zeros / truncates before transforming
            short = x[:, : n - 3]
            padded = np.concatenate([np.asarray(short), np.zeros((5, 3), dtype=short.dtype)], axis=1)
            assert eq(np.fft.fft(short, n=n), F.fft(padded, omega))
            if x.dtype.itemsize >= 4:
                dx = x.to_device()
                assert eq(gb.ntt_device(dx).cpu(), want)
                assert eq(gb.ntt_device(gb.ntt_device(dx), inverse=True).cpu(), x)
    with pytest.raises(ValueError):
        np.fft.fft(gb.GF(31).Random(7, seed=1))  # 7 does not divide 30


# ------------------------------------------------------------------------------------------------ dense linear algebra
@pytest.mark.parametrize("name", FIELDS)
def test_ref_vectors_linalg(ref_linalg, name):
    """tests/fields/test_linalg.py:301-550 vs Sage, first and last dtype of each field."""
    d = ref_linalg
    GF = make_gf(name)

    def cases(op):
        return range(int(d[f"{name}/{op}/count"]))

    for dtype in (GF.dtypes[0], GF.dtypes[-1]):
        for i in cases("row_reduce"):
            z = GF(d[f"{name}/row_reduce/X{i}"], dtype=dtype).row_reduce()
            assert type(z) is GF and z.dtype == dtype and eq(z, d[f"{name}/row_reduce/Z{i}"])
        for i in cases("lu_decompose"):
            L, U = GF(d[f"{name}/lu_decompose/X{i}"], dtype=dtype).lu_decompose()
            assert type(L) is GF and type(U) is GF
            assert eq(L, d[f"{name}/lu_decompose/L{i}"]) and eq(U, d[f"{name}/lu_decompose/U{i}"])
        for i in cases("plu_decompose"):
            P, L, U = GF(d[f"{name}/plu_decompose/X{i}"], dtype=dtype).plu_decompose()
            assert type(P) is GF and type(L) is GF and type(U) is GF
            assert eq(P, d[f"{name}/plu_decompose/P{i}"])
            assert eq(L, d[f"{name}/plu_decompose/L{i}"])
            assert eq(U, d[f"{name}/plu_decompose/U{i}"])
        for i in cases("matrix_inverse"):
            z = np.linalg.inv(GF(d[f"{name}/matrix_inverse/X{i}"], dtype=dtype))
            assert type(z) is GF and z.dtype == dtype and eq(z, d[f"{name}/matrix_inverse/Z{i}"])
        for i in cases("matrix_determinant"):
            z = np.linalg.det(GF(d[f"{name}/matrix_determinant/X{i}"], dtype=dtype))
            assert type(z) is GF and int(z) == int(d[f"{name}/matrix_determinant/Z{i}"])
        for i in cases("matrix_solve"):
            z = np.linalg.solve(GF(d[f"{name}/matrix_solve/X{i}"], dtype=dtype), GF(d[f"{name}/matrix_solve/Y{i}"], dtype=dtype))
            assert type(z) is GF and eq(z, d[f"{name}/matrix_solve/Z{i}"])
        for op in ("row_space", "column_space", "left_null_space", "null_space"):
            for i in cases(op):
                want = d[f"{name}/{op}/Z{i}"]
                got = getattr(GF(d[f"{name}/{op}/X{i}"], dtype=dtype), op)()
                assert type(got) is GF and got.shape[0] == want.shape[0]
                assert np.array_equal(np.asarray(got).astype(np.int64).reshape(want.shape), want), (name, op, i)


def test_live_linalg_next(live_next):
    data, meta = live_next
    for fname, m in meta["linalg"].items():
        GF = live_gf(m)
        for case in m["cases"]:
            b = f"linalg/{fname}/{case}"
            A = GF(data[f"{b}/A"])
            assert eq(A.row_reduce(), data[f"{b}/rre"]), b
            assert eq(A.row_reduce(ncols=max(1, A.shape[1] // 2)), data[f"{b}/rre_ncols"]), b
            assert eq(A.row_reduce(eye="right"), data[f"{b}/rre_right"]), b
            assert np.linalg.matrix_rank(A) == int(data[f"{b}/rank"])
            P, L, U = A.plu_decompose()
            assert eq(P, data[f"{b}/P"]) and eq(L, data[f"{b}/L"]) and eq(U, data[f"{b}/U"]), b
            assert eq(P @ L @ U, A)
            for op in ("row_space", "column_space", "left_null_space", "null_space"):
                want = data[f"{b}/{op}"]
                got = getattr(A, op)()
                assert got.shape[0] == want.shape[0]
                assert np.array_equal(np.asarray(got).astype(np.int64).reshape(want.shape), want), (b, op)
            if f"{b}/lu_L" in data:
                L2, U2 = A.lu_decompose()
                assert eq(L2, data[f"{b}/lu_L"]) and eq(U2, data[f"{b}/lu_U"]), b
            elif A.shape[0] - 1 <= A.shape[1]:
                with pytest.raises(ValueError):
                    A.lu_decompose()
            else:  # more rows than the diagonal walk has columns for: the reference fails one way or the other
                with pytest.raises((ValueError, IndexError)):
                    A.lu_decompose()
            if f"{b}/det" in data:
                assert int(np.linalg.det(A)) == int(data[f"{b}/det"]), b
            if f"{b}/inv" in data:
                assert eq(np.linalg.inv(A), data[f"{b}/inv"]), b
                assert eq(np.linalg.solve(A, GF(data[f"{b}/b1"])), data[f"{b}/x1"]), b
                assert eq(np.linalg.solve(A, GF(data[f"{b}/b2"])), data[f"{b}/x2"]), b
            elif f"{b}/inv_raises" in data:
                with pytest.raises(np.linalg.LinAlgError) as ei:
                    np.linalg.inv(A)
                assert str(ei.value) == str(data[f"{b}/inv_raises"]), b


def test_linalg_vs_oracle_larger():
    """Sizes past the fixtures: several elimination tiles, every storage width, vs the oracle and A A^-1 = I."""
    for GF, F, n in [
        (gb.GF(998244353), oracle.Field.prime(998244353), 300),
        (gb.GF(2**8), None, 260),
        (gb.GF(2**31 - 1), oracle.Field.prime(2**31 - 1), 150),
        (gb.GF(2**20), None, 100),
    ]:
        if F is None:
            m = GF.degree
            F = oracle.Field.binary(m, int(GF.irreducible_poly), alpha=int(GF.primitive_element) if m <= 16 else None)
        A = GF.Random((n, n), seed=n)
        Ai = np.linalg.inv(A)
        assert eq(Ai, F.matrix_inv(np.asarray(A)))
        assert eq(A @ Ai, np.identity(n))
        assert int(np.linalg.det(A)) == F.det(np.asarray(A))
        W = GF.Random((n // 2, n + 37), seed=n + 1)
        R, p = F.row_reduce(np.asarray(W), ncols=n // 3)
        assert eq(W.row_reduce(ncols=n // 3), R)
        N = W.null_space()
        assert N.shape[0] == W.shape[1] - np.linalg.matrix_rank(W)
        assert not np.asarray(W @ N.T).any()
        b = GF.Random(n, seed=n + 2)
        x = np.linalg.solve(A, b)
        assert eq(A @ x, b)
    # a singular matrix: determinant 0, LinAlgError with the reference's message
    GF = gb.GF(31)
    S = GF.Random((6, 3), seed=1) @ GF.Random((3, 6), seed=2)
    assert int(np.linalg.det(S)) == 0
    with pytest.raises(np.linalg.LinAlgError, match="does not have full rank of 6, but rank of 3"):
        np.linalg.inv(S)


def test_dot_family_matches_oracle():
    GF = gb.GF(2**8)
    F = oracle.Field.binary(8, 0x11D, alpha=2)
    a, b = GF.Random(50, seed=1), GF.Random(50, seed=2)
    want = 0
    for v in F.mul(np.asarray(a), np.asarray(b)):
        want ^= int(v)
    assert int(np.dot(a, b)) == want and int(np.vdot(a, b)) == want and int(np.inner(a, b)) == want
    assert eq(np.outer(a, b), F.mul(np.asarray(a)[:, None], np.asarray(b)[None, :]))
    A, B = GF.Random((9, 50), seed=3), GF.Random((50, 4), seed=4)
    assert eq(np.dot(A, B), F.matmul(np.asarray(A), np.asarray(B)))
    assert eq(np.dot(A, a), F.matmul(np.asarray(A), np.asarray(a).reshape(-1, 1)).reshape(-1))


def test_fft_mixed_host_pipeline_many_chunks():
    """A host batch larger than one staging chunk (32 MiB): every chunk gets its own, smaller workspace layout."""
    GF = gb.GF(998244353)
    F = oracle.Field.prime(998244353)
    n = 7 * 17 * 2**13
    x = GF.Random((19, n), seed=5)
    X = np.fft.fft(x)
    omega = int(GF.primitive_root_of_unity(n))
    for row in (0, 7, 8, 18):
        assert eq(X[row], F.fft(np.asarray(x[row]), omega)), row
    assert eq(np.fft.ifft(X), x)


def test_row_reduce_panel_edges():
    """Shapes around the 32-column panel and 64-row tile boundaries, rank-deficient panels, device-resident entry."""
    from galois_b200 import _linalg

    rng = np.random.default_rng(8)
    GF = gb.GF(3191)
    F = oracle.Field.prime(3191)
    for rows, cols, ncols in [(31, 33, None), (32, 32, None), (33, 95, 40), (65, 64, None), (100, 130, 64), (70, 200, 33),
                              (5, 300, None), (300, 5, None), (129, 129, 128)]:
        A = rng.integers(0, 3191, (rows, cols))
        A[:, rng.integers(0, cols, 3)] = 0                # zero columns: null steps inside a panel
        A[rng.integers(0, rows, 2), :] = 0                # zero rows
        if cols > 40:
            A[:, 35] = A[:, 2]                            # a dependent column
        R, p = F.row_reduce(A, ncols)
        got = GF(A).row_reduce(ncols=ncols)
        assert eq(got, R), (rows, cols, ncols)
        dA = GF(A).to_device()
        out, info, _ = _linalg.row_reduce_device(dA, ncols=ncols)
        assert eq(out.cpu(), R) and int(info[0].item()) == p
    # sparse 0/1 matrices over GF(2): many exchanges and null steps
    G2 = gb.GF(2)
    F2 = oracle.Field.prime(2)
    for rows, cols in [(64, 200), (130, 70)]:
        A = (rng.random((rows, cols)) < 0.05).astype(np.int64)
        assert eq(G2(A).row_reduce(), F2.row_reduce(A)[0])
        assert eq(G2(A).left_null_space(), F2.left_null_space(A))


def test_fft_hybrid_odd_times_power_of_two():
    """N = n_odd * 2^k with k >= 5: odd radices on columns, then every power-of-two kernel family (single kernel,
    short-column and tile four-step) on rows, then the output permutation; forward, inverse, 4- and 8-byte elements."""
    GF = gb.GF(998244353)
    F = oracle.Field.prime(998244353)
    for n, batch in [(7 * 32, 5), (17 * 64, 3), (119 * 32, 2), (7 * 1024, 3), (7 * 2048, 3), (119 * 4096, 2), (17 * 2**15, 2),
                     (7 * 2**17, 2)]:
        omega = int(GF.primitive_root_of_unity(n))
        x = GF.Random((batch, n), seed=n)
        want = F.fft(np.asarray(x), omega, threads=8)
        X = np.fft.fft(x)
        assert eq(X, want), n
        assert eq(np.fft.ifft(X), x), n
        assert eq(np.fft.fft(x.astype(np.int64)), want), n
        dx = x.to_device()
        assert eq(gb.ntt_device(dx).cpu(), want), n
    # a large batch of small transforms (row chunks larger than one grid dimension)
    x = GF.Random((70000, 224), seed=3)
    X = np.fft.fft(x)
    omega = int(GF.primitive_root_of_unity(224))
    for row in (0, 32767, 32768, 69999):
        assert eq(X[row], F.fft(np.asarray(x[row]), omega)), row


def test_row_reduce_slab_in_global_memory():
    """More rows than the shared-memory slab holds (8-byte elements: > 775 rows): the panel works on global memory."""
    GF = gb.GF(998244353)
    F = oracle.Field.prime(998244353)
    A = GF.Random((800, 840), seed=77, dtype=np.int64)
    R, p = F.row_reduce(np.asarray(A), ncols=820)
    got = A.row_reduce(ncols=820)
    assert got.dtype == np.int64 and eq(got, R)
    S = GF.Random((790, 790), seed=78, dtype=np.int64)
    Si = np.linalg.inv(S)
    assert eq(S @ Si, np.identity(790))
    G8 = gb.GF(2**8)
    B = G8.Random((6300, 40), seed=79)  # one-byte elements: 6300 rows exceed the slab as well
    F8 = oracle.Field.binary(8, 0x11D, alpha=2)
    assert eq(B.row_reduce(), F8.row_reduce(np.asarray(B))[0])


@pytest.mark.parametrize("p,na,nb", [(998244353, 4096, 4096), (998244353, 1000, 33), (998244353, 5000, 7), (998244353, 30000, 20000),
                                     (65537, 30000, 30000), (65537, 35000, 31000), (2**31 - 1, 3000, 2000), (7681, 300, 200),
                                     (3329, 2000, 1500)])
def test_convolve_ntt_route_matches_the_oracle(p, na, nb):
    """np.convolve over GF(p): the NTT route (padded length 2^k | p - 1) and the schoolbook kernel give the reference's
    field result (convolve_jit, _domains/_function.py:111-167), every storage width."""
    import oracle

    GF = gb.GF(p)
    F = oracle.Field.prime(p)
    rng = np.random.default_rng(na + nb)
    a = rng.integers(0, p, na)
    b = rng.integers(0, p, nb)
    a[:3] = [p - 1, 0, 1]
    want = F.convolve(a, b)
    for dtype in GF.dtypes[:1] + GF.dtypes[-1:]:
        got = np.convolve(GF(a, dtype=dtype), GF(b, dtype=dtype))
        assert type(got) is GF and got.dtype == dtype
        assert np.array_equal(np.asarray(got).astype(np.int64), want), (p, na, nb, dtype)
    got = np.convolve(GF(b), GF(a))
    assert np.array_equal(np.asarray(got).astype(np.int64), want)
