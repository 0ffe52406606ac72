"""
Randomised cross-check of the oracle against the LIVE reference for the §8(f) rows (build container only: needs
/root/reference).  Not part of the pytest suites; run by hand when the oracle changes:

    NUMBA_CACHE_DIR=/tmp/numba python tests/golden/fuzz_oracle_vs_reference.py [cases]

Covers shapes / ranks / field kinds the committed fixtures do not enumerate: row_reduce (ncols, eye), LU / PLU, inverse,
determinant, solve, rank, the four subspaces, polynomial roots and mixed-radix DFTs.
"""

from __future__ import annotations

import os
import sys
import warnings
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
warnings.filterwarnings("ignore")
import galois  # noqa: E402

import oracle  # noqa: E402


def fields():
    out = []
    for q in (2, 3, 31, 3191, 998244353, 2**31 - 1):
        out.append((galois.GF(q), oracle.Field.prime(q)))
    for m in (2, 4, 8, 12):
        G = galois.GF(2**m)
        out.append((G, oracle.Field.binary(m, int(G.irreducible_poly), alpha=int(G.primitive_element))))
    G = galois.GF(2**20)
    out.append((G, oracle.Field.binary(20, int(G.irreducible_poly))))
    return out


def main(cases: int) -> int:
    rng = np.random.default_rng(12345)
    n_checked = 0
    for GF, F in fields():
        q = GF.order
        for _ in range(cases):
            m, n = int(rng.integers(1, 9)), int(rng.integers(1, 9))
            A = rng.integers(0, q, (m, n))
            if rng.random() < 0.5:  # lower the rank / add structure
                r = int(rng.integers(0, min(m, n) + 1))
                A = np.asarray(GF(rng.integers(0, q, (m, r))) @ GF(rng.integers(0, q, (r, n)))) if r else np.zeros((m, n), dtype=np.int64)
            if rng.random() < 0.3:
                A[:, int(rng.integers(0, n))] = 0
            Ag = GF(A)
            ncols = int(rng.integers(1, n + 1))
            assert np.array_equal(F.row_reduce(A, ncols)[0], np.asarray(Ag.row_reduce(ncols=ncols))), ("rre", GF.name, A)
            assert np.array_equal(F.row_reduce_eye(A, eye="right"), np.asarray(Ag.row_reduce(eye="right")))
            assert F.matrix_rank(A) == int(np.linalg.matrix_rank(Ag))
            P, L, U, _ = F.plu_decompose(A)
            Pg, Lg, Ug = Ag.plu_decompose()
            assert np.array_equal(P, np.asarray(Pg)) and np.array_equal(L, np.asarray(Lg)) and np.array_equal(U, np.asarray(Ug)), ("plu", GF.name, A)
            try:
                Lr, Ur = Ag.lu_decompose()
                L2, U2 = F.lu_decompose(A)
                assert np.array_equal(L2, np.asarray(Lr)) and np.array_equal(U2, np.asarray(Ur))
            except (ValueError, IndexError):
                if m - 1 <= n:
                    try:
                        F.lu_decompose(A)
                        raise AssertionError(("lu should fail", GF.name, A))
                    except ValueError:
                        pass
            for op in ("row_space", "column_space", "left_null_space", "null_space"):
                want = np.asarray(getattr(Ag, op)())
                got = getattr(F, op)(A)
                assert got.shape[0] == want.shape[0] and np.array_equal(got.reshape(want.shape), want), (op, GF.name, A)
            if m == n:
                assert F.det(A) == int(np.linalg.det(Ag)), ("det", GF.name, A)
                try:
                    inv = np.asarray(np.linalg.inv(Ag))
                    assert np.array_equal(F.matrix_inv(A), inv)
                    b = rng.integers(0, q, (n, 2))
                    assert np.array_equal(F.solve(A, b), np.asarray(np.linalg.solve(Ag, GF(b))))
                except np.linalg.LinAlgError as e:
                    try:
                        F.matrix_inv(A)
                        raise AssertionError("inverse should fail")
                    except np.linalg.LinAlgError as e2:
                        assert str(e) == str(e2)
            n_checked += 1
        # polynomial roots (fields small enough for the reference's serial search)
        if q <= 2**12:
            for _ in range(cases):
                deg = int(rng.integers(1, 12))
                c = rng.integers(0, q, deg + 1)
                c[0] = int(rng.integers(1, q)) if q > 2 else 1
                if rng.random() < 0.5:
                    c[int(rng.integers(1, deg + 1))] = 0
                poly = galois.Poly(GF(c))
                want = np.asarray(poly.roots())
                got = F.poly_roots(poly.nonzero_degrees, np.asarray(poly.nonzero_coeffs), primitive_element=int(GF.primitive_element))
                assert np.array_equal(got, want), ("roots", GF.name, c)
                n_checked += 1
        # mixed-radix DFT at a few divisors of q - 1
        divs = [d for d in range(2, 200) if (q - 1) % d == 0][:6]
        for n in divs:
            x = GF.Random(n, seed=int(rng.integers(1 << 30)))
            omega = int(GF.primitive_root_of_unity(n))
            assert np.array_equal(F.fft(np.asarray(x), omega), np.asarray(np.fft.fft(x))), ("fft", GF.name, n)
            n_checked += 1
        print(f"{GF.name}: ok", flush=True)
    print(f"{n_checked} cases agree")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 60))
