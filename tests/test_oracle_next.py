"""
Pins the oracle's restatement of the SURVEY.md §8(f) rows — Poly.roots (Chien search), the mixed-radix DFT and dense
linear algebra (row_reduce / LU / PLU / inv / det / solve / rank / row, column and null spaces) — against
  * the reference's Sage vectors (ref_linalg.npz, ref_roots.npz; /root/reference/tests/fields/data, tests/polys/data),
  * outputs of the live reference (live_next.npz; tests/golden/make_golden.py).
No GPU needed.
"""

import numpy as np
import pytest

import oracle
from helpers import oracle_field, parse_field

FIELDS = ["GF(2)", "GF(2^2)", "GF(2^3)", "GF(2^8)", "GF(2^32)", "GF(5)", "GF(7)", "GF(31)", "GF(3191)", "GF(2147483647)"]


def live_field(meta: dict):
    """oracle field from a live_next_meta.json entry."""
    q = meta["order"]
    if q & (q - 1) == 0 and q > 1 and meta["irreducible_poly"] >= 2 and q != 2:
        m = q.bit_length() - 1
        return oracle.Field.binary(m, meta["irreducible_poly"], alpha=meta["primitive_element"] if m <= 16 else None)
    if q == 2:
        return oracle.Field.prime(2)
    return oracle.Field.prime(q)


def _alpha(name):
    import json
    from helpers import GOLDEN

    meta = json.load(open(GOLDEN / "ref_fields_meta.json"))
    return int(meta[name]["primitive_element"])


@pytest.mark.parametrize("name", FIELDS)
def test_ref_vectors_linalg(ref_linalg, name):
    """tests/fields/test_linalg.py:301-550 vs Sage."""
    d = ref_linalg
    F = oracle_field(name)

    def cases(op):
        return range(int(d[f"{name}/{op}/count"]))

    for i in cases("row_reduce"):
        assert np.array_equal(F.row_reduce(d[f"{name}/row_reduce/X{i}"])[0], d[f"{name}/row_reduce/Z{i}"])
    for i in cases("lu_decompose"):
        L, U = F.lu_decompose(d[f"{name}/lu_decompose/X{i}"])
        assert np.array_equal(L, d[f"{name}/lu_decompose/L{i}"]) and np.array_equal(U, d[f"{name}/lu_decompose/U{i}"])
    for i in cases("plu_decompose"):
        P, L, U, _ = F.plu_decompose(d[f"{name}/plu_decompose/X{i}"])
        assert np.array_equal(P, d[f"{name}/plu_decompose/P{i}"])
        assert np.array_equal(L, d[f"{name}/plu_decompose/L{i}"])
        assert np.array_equal(U, d[f"{name}/plu_decompose/U{i}"])
    for i in cases("matrix_inverse"):
        assert np.array_equal(F.matrix_inv(d[f"{name}/matrix_inverse/X{i}"]), d[f"{name}/matrix_inverse/Z{i}"])
    for i in cases("matrix_determinant"):
        assert F.det(d[f"{name}/matrix_determinant/X{i}"]) == int(d[f"{name}/matrix_determinant/Z{i}"])
    for i in cases("matrix_solve"):
        x = F.solve(d[f"{name}/matrix_solve/X{i}"], d[f"{name}/matrix_solve/Y{i}"])
        assert np.array_equal(x, d[f"{name}/matrix_solve/Z{i}"])
    for op, fn in [("row_space", F.row_space), ("column_space", F.column_space),
                   ("left_null_space", F.left_null_space), ("null_space", F.null_space)]:
        for i in cases(op):
            z = d[f"{name}/{op}/Z{i}"]
            got = fn(d[f"{name}/{op}/X{i}"])
            assert got.shape[0] == z.shape[0] and np.array_equal(got.reshape(z.shape), z), (name, op, i)


@pytest.mark.parametrize("name", FIELDS)
def test_ref_vectors_roots(ref_roots, name):
    """tests/polys/test_arithmetic_methods.py:28-43 vs Sage: roots in increasing order."""
    d = ref_roots
    F = oracle_field(name)
    alpha = _alpha(name)
    if name in ("GF(2^32)", "GF(2147483647)"):
        # Chien search over >= 2^31 elements: minutes per polynomial on the CPU.  These Sage vectors are checked
        # directly against the CUDA path in tests/test_gpu_next.py instead.
        pytest.skip("exhaustive search over >= 2^31 elements is too slow for the CPU suite")
    for i in range(int(d[f"{name}/count"])):
        coeffs = d[f"{name}/X{i}"]
        nz = np.nonzero(coeffs)[0]
        if nz.size == 0:
            continue
        degrees = (coeffs.size - 1 - nz).astype(np.int64)
        r = F.poly_roots(degrees, coeffs[nz], primitive_element=alpha)
        assert np.array_equal(r, d[f"{name}/R{i}"]), (name, i)


def test_live_roots(live_next):
    data, meta = live_next
    for name in data["roots/names"]:
        fname = str(name).split("/")[0]
        F = live_field(meta["roots"][fname])
        r = F.poly_roots(data[f"roots/{name}/degrees"], data[f"roots/{name}/coeffs"],
                         primitive_element=meta["roots"][fname]["primitive_element"])
        assert np.array_equal(r, data[f"roots/{name}/roots"]), name


def test_live_fft_mixed_radix(live_next):
    data, meta = live_next
    for fname, m in meta["fft"].items():
        F = live_field(m)
        for n, omega in zip(m["sizes"], m["omegas"]):
            x = data[f"fft/{fname}/{n}/x"]
            assert np.array_equal(F.fft(x, omega), data[f"fft/{fname}/{n}/X"]), (fname, n)
            omega_inv = int(F.inv(np.array([omega]))[0])
            p = 2 if m["order"] & (m["order"] - 1) == 0 else m["order"]
            scale = int(F.inv(np.array([n % p]))[0])
            assert np.array_equal(F.fft(x, omega_inv, scale=scale), data[f"fft/{fname}/{n}/ix"]), (fname, n)


def test_live_linalg_next(live_next):
    data, meta = live_next
    for fname, m in meta["linalg"].items():
        F = live_field(m)
        for case in m["cases"]:
            b = f"linalg/{fname}/{case}"
            A = data[f"{b}/A"]
            assert np.array_equal(F.row_reduce(A)[0], data[f"{b}/rre"]), b
            assert np.array_equal(F.row_reduce(A, ncols=max(1, A.shape[1] // 2))[0], data[f"{b}/rre_ncols"]), b
            assert np.array_equal(F.row_reduce_eye(A, eye="right"), data[f"{b}/rre_right"]), b
            assert F.matrix_rank(A) == int(data[f"{b}/rank"])
            P, L, U, _ = F.plu_decompose(A)
            assert np.array_equal(P, data[f"{b}/P"]) and np.array_equal(L, data[f"{b}/L"]) and np.array_equal(U, data[f"{b}/U"]), b
            for op in ("row_space", "column_space", "left_null_space", "null_space"):
                z = data[f"{b}/{op}"]
                got = getattr(F, op)(A)
                assert got.shape[0] == z.shape[0] and np.array_equal(got.reshape(z.shape), z), (b, op)
            if f"{b}/lu_L" in data:
                L2, U2 = F.lu_decompose(A)
                assert np.array_equal(L2, data[f"{b}/lu_L"]) and np.array_equal(U2, data[f"{b}/lu_U"]), b
            elif A.shape[0] <= A.shape[1]:
                with pytest.raises(ValueError):
                    F.lu_decompose(A)
            if f"{b}/det" in data:
                assert F.det(A) == int(data[f"{b}/det"]), b
            if f"{b}/inv" in data:
                assert np.array_equal(F.matrix_inv(A), data[f"{b}/inv"]), b
                assert np.array_equal(F.solve(A, data[f"{b}/b1"]), data[f"{b}/x1"]), b
                assert np.array_equal(F.solve(A, data[f"{b}/b2"]), data[f"{b}/x2"]), b
            elif f"{b}/inv_raises" in data:
                with pytest.raises(np.linalg.LinAlgError) as ei:
                    F.matrix_inv(A)
                assert str(ei.value) == str(data[f"{b}/inv_raises"]), b


def test_oracle_prime_field_above_int64_products():
    """p > ~3.03e9: a * b exceeds 2^63 (the reference leaves int64 for Python integers there, _fields/_factory.py dtypes).  The
    oracle multiplies in 128 bits; checked against Python big integers, element-wise and through the elimination."""
    p = 4294967291
    F = oracle.Field.prime(p)
    rng = np.random.default_rng(11)
    a = rng.integers(0, p, 2000)
    b = rng.integers(1, p, 2000)
    assert [int(x) for x in F.mul(a, b)] == [int(x) * int(y) % p for x, y in zip(a, b)]
    assert [int(x) for x in F.mul(F.div(a, b), b)] == [int(x) for x in a]
    A = rng.integers(0, p, (7, 40))
    A[6] = A[0]
    R, n_piv = F.row_reduce(A, 33)
    assert n_piv == 6 and int(R.max()) < p and int(R.min()) >= 0
    # the reduced rows span the same space: every original row is a combination of them (checked with big integers)
    piv_cols = [int(np.flatnonzero(R[i])[0]) for i in range(n_piv)]
    for row in A:
        acc = [0] * 40
        for i, c in enumerate(piv_cols):
            f = int(row[c])
            for j in range(40):
                acc[j] = (acc[j] + f * int(R[i, j])) % p
        assert acc == [int(x) for x in row]
