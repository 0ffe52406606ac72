"""The cluster panel of the Gaussian elimination (k_rr_panel_cluster + k_rr_panel_w2, gfb_solve.cu) against the single-CTA
panel (GFB_ELIM_CLUSTER=0) and the oracle: same reduced rows, pivot count and determinant, bit for bit, over the fields whose
elements fit 32 bits (prime fields below 2^32, GF(2), GF(2^m) with tables) — row_reduce_jit (_domains/_linalg.py:321-353)."""

import os

import numpy as np
import pytest

import galois_b200 as gb
import oracle
from galois_b200 import _linalg

pytestmark = pytest.mark.gpu


def _both(field, A, ncols=None):
    """(R, n_pivots, det) through the cluster panel and through the single-CTA panel."""
    out = []
    for flag in ("1", "0"):
        os.environ["GFB_ELIM_CLUSTER"] = flag
        try:
            out.append(_linalg._row_reduce_raw(field, A, ncols, want_det=True))
        finally:
            os.environ.pop("GFB_ELIM_CLUSTER", None)
    return out


CASES = [(1, 1, None), (2, 3, None), (3, 3, None), (7, 40, 33), (31, 33, None), (32, 32, None), (33, 95, 40), (64, 64, None),
         (65, 64, None), (100, 130, 64), (5, 300, None), (300, 5, None), (129, 129, 128), (257, 300, 290)]


@pytest.mark.parametrize("order", [998244353, 2**31 - 1, 4294967291, 7, 2, 2**8, 2**4])
def test_cluster_panel_matches_single_cta_and_oracle(order):
    GF = gb.GF(order)
    if order in (2**8, 2**4):
        m = order.bit_length() - 1
        F = oracle.Field.binary(m, int(GF.irreducible_poly), alpha=int(GF.primitive_element))
    else:
        F = oracle.Field.prime(order)
    rng = np.random.default_rng(order % 1000)
    for rows, cols, ncols in CASES:
        A = rng.integers(0, order, (rows, cols))
        if cols > 4:
            A[:, rng.integers(0, cols, 2)] = 0            # columns without a pivot inside a panel
        if rows > 4:
            A[rng.integers(0, rows, 2), :] = 0            # zero rows
            A[rows - 1] = A[0]                            # a dependent row
        if cols > 40:
            A[:, 35] = A[:, 2]                            # a dependent column
        if rows > 8:
            A[: rows // 2, 0] = 0                         # the first pivot comes from the lower half: an exchange
        (R1, p1, d1), (R0, p0, d0) = _both(GF, GF(A), ncols)
        Ro, po = F.row_reduce(A, ncols)
        assert p1 == p0 == po, (order, rows, cols, ncols)
        assert np.array_equal(np.asarray(R1), np.asarray(R0)), (order, rows, cols, ncols)
        assert np.array_equal(np.asarray(R1).astype(np.int64), np.asarray(Ro).astype(np.int64)), (order, rows, cols, ncols)
        assert int(d1[0]) == int(d0[0]), (order, rows, cols, ncols)


def test_cluster_panel_full_rank_determinant_and_inverse():
    GF = gb.GF(998244353)
    F = oracle.Field.prime(998244353)
    for n in (5, 32, 33, 100, 256):
        A = GF.Random((n, n), seed=n)
        (R1, p1, d1), (R0, p0, d0) = _both(GF, A)
        assert p1 == p0 and int(d1[0]) == int(d0[0])
        assert int(np.linalg.det(A)) == int(F.det(np.asarray(A)))
        Ai = np.linalg.inv(A)
        assert np.array_equal(np.asarray(A @ Ai), np.identity(n, dtype=np.asarray(A).dtype))


def test_cluster_panel_several_rows_per_thread():
    """1024 < rows <= 4096: four rows per thread of the cluster."""
    GF = gb.GF(2**31 - 1)
    F = oracle.Field.prime(2**31 - 1)
    A = GF.Random((1500, 70), seed=5)
    A[:, 40] = A[:, 3]
    A[700] = A[2]
    (R1, p1, d1), (R0, p0, d0) = _both(GF, A, 66)
    assert p1 == p0 and np.array_equal(np.asarray(R1), np.asarray(R0))
    Ro, po = F.row_reduce(np.asarray(A), 66)
    assert po == p1 and np.array_equal(np.asarray(R1).astype(np.int64), np.asarray(Ro).astype(np.int64))
    G2 = gb.GF(2)
    B = G2((np.random.default_rng(3).random((3000, 90)) < 0.02).astype(np.int64))
    (R1, p1, _), (R0, p0, _) = _both(G2, B)
    assert p1 == p0 and np.array_equal(np.asarray(R1), np.asarray(R0))


def test_cluster_panel_1024_inverse_round_trip():
    GF = gb.GF(998244353)
    A = GF.Random((1024, 1024), seed=2024)
    Ai = np.linalg.inv(A)
    assert np.array_equal(np.asarray(A @ Ai), np.identity(1024, dtype=np.asarray(A).dtype))
