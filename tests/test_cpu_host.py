"""
CPU-only tests (``-m "not gpu"``): host logic of the mirror (field construction, number theory, argument
checking) and that the C-ABI library loads and exports every symbol include/galois_b200.h declares.
No compute call is made: without a GPU the product path must fail loudly, never fall back.
"""

import re
from pathlib import Path

import numpy as np
import pytest

import galois_b200 as gb
from galois_b200 import _lib, _nt

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "galois_b200.h").read_text()
    declared = set(re.findall(r"\b(gfb_[a-z0-9_]+)\s*\(", header))
    declared -= {"gfb_field", "gfb_ntt_plan", "gfb_rs_code"}
    assert len(declared) >= 30
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libgalois_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.gfb_version()


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    GF = gb.GF(31)
    x = GF([1, 2, 3])
    with pytest.raises(gb.BackendUnavailable):
        x * x
    with pytest.raises(gb.BackendUnavailable):
        gb.ntt([1, 2, 3, 4], modulus=5)


def test_product_never_imports_oracle():
    for path in (ROOT / "galois_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), path
    for path in (ROOT / "galois_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(), path


def test_number_theory_matches_reference_constants(live_ntt, live_rs, live_elementwise):
    _, meta = live_ntt
    for key, m in meta.items():
        if "seed" in key:
            continue
        p = int(key.split("/")[0][1:])
        n = int(key.split("/")[1][1:])
        assert _nt.primitive_root(p) == m["primitive_element"]
        if n > 1:
            assert int(gb.GF(p).primitive_root_of_unity(n)) == m["omega"]
    assert _nt.primitive_root(2**31 - 1) == 7  # tests/fields/data/GF(2147483647)/properties.json
    assert _nt.primitive_root(998244353) == 3
    _, emeta = live_elementwise
    for name, m in emeta.items():
        deg = int(re.match(r"GF\(2\^(\d+)\)", name).group(1))
        assert _nt.CONWAY_2[deg] == m["irreducible_poly"]
        assert m["primitive_element"] == 2
        assert _nt.is_irreducible2(_nt.CONWAY_2[deg])
    assert _nt.is_primitive_element2(2, _nt.CONWAY_2[16])
    assert not _nt.is_irreducible2(0b10001)  # x^4 + 1
    assert _nt.factors(2**31 - 2) == ((2, 3, 7, 11, 31, 151, 331), (1, 2, 1, 1, 1, 1, 1))
    assert _nt.prime_factors_with_multiplicity(176) == [2, 2, 2, 2, 11]
    assert _nt.is_prime(998244353) and not _nt.is_prime(998244351)
    assert _nt.parse_poly2("x^8 + x^4 + x^3 + x^2 + 1") == 0x11D


def test_field_factory_and_array_construction():
    GF = gb.GF(2**31 - 1)
    assert gb.GF(2**31 - 1) is GF  # flyweight (_fields/_factory.py:385)
    assert GF.characteristic == 2**31 - 1 and GF.degree == 1 and GF.order == 2**31 - 1
    assert [np.dtype(d).name for d in GF.dtypes] == ["uint32", "int32", "int64"]  # _fields/_ufunc.py:36-48
    assert int(GF.primitive_element) == 7
    x = GF.Random(10, seed=1)
    assert type(x) is GF and x.dtype == np.uint32
    want = np.random.default_rng(1).integers(0, GF.order, 10, dtype=np.uint32)  # _domains/_array.py:278-280
    assert np.array_equal(np.asarray(x), want)
    assert np.array_equal(np.asarray(GF.Zeros(3)), [0, 0, 0]) and np.array_equal(np.asarray(GF.Ones(2)), [1, 1])
    with pytest.raises(ValueError):
        GF([GF.order])
    with pytest.raises(TypeError):
        GF([1.5])
    with pytest.raises(TypeError):
        GF([1, 2], dtype=np.uint8)
    G8 = gb.GF(2**8)
    assert G8.irreducible_poly == 0x11D and int(G8.primitive_element) == 2
    assert [np.dtype(d).name for d in G8.dtypes] == ["uint8", "uint16", "uint32", "int16", "int32", "int64"]
    with pytest.raises(LookupError):
        gb.GF(2**128)  # no Conway polynomial of degree 128, as in the reference
    G128 = gb.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
    assert G128.dtypes == [np.object_]
    v = G128([1 << 127, 3])
    assert np.array_equal(G128.from_limbs(v.to_limbs()), v)
    with pytest.raises(ValueError):
        gb.GF(2**4, irreducible_poly=0b10001)
    with pytest.raises(ValueError):
        gb.GF(6)
    with pytest.raises(NotImplementedError):
        gb.GF(3**4)
    # comparisons / views behave like ndarrays
    a = GF([1, 2, 3])
    assert (a == GF([1, 5, 3])).tolist() == [True, False, True]
    assert type(a[1:]) is GF and type(a.reshape(3, 1)) is GF


def test_reed_solomon_constructor_checks():
    rs = gb.ReedSolomon(255, 223)
    assert (rs.n, rs.k, rs.d, rs.t) == (255, 223, 33, 16)
    assert rs.field.irreducible_poly == 0x11D and int(rs.alpha) == 2 and rs.c == 1
    assert rs.is_primitive and rs.is_narrow_sense and rs.is_systematic
    assert gb.ReedSolomon(255, d=33).k == 223
    assert gb.ReedSolomon(15, 9).field.order == 16
    with pytest.raises(ValueError):
        gb.ReedSolomon(255, 223, 30)
    with pytest.raises(ValueError):
        gb.ReedSolomon(255)
    with pytest.raises(TypeError):
        gb.ReedSolomon(255.0, 223)
    with pytest.raises(ValueError):
        gb.ReedSolomon(255, 223, c=-1)
    assert np.array_equal(np.asarray(rs.roots)[:4], [2, 4, 8, 16])


def test_ntt_argument_checks_need_no_gpu():
    with pytest.raises(TypeError):
        gb.ntt(1)
    with pytest.raises(ValueError):
        gb.ntt([1, 2, 3, 4], size=3)
    with pytest.raises(ValueError):
        gb.ntt([1, 2, 3, 40], modulus=43)
    with pytest.raises(ValueError):
        gb.ntt([1, 2, 3, 4], modulus=3 * 256 + 2)


def test_bch_construction_matches_reference(live_bch):
    """BCH.__init__ host logic (generator polynomial from d / search from k, _codes/_bch.py:1178-1252) vs the live
    reference's codes; no GPU needed."""
    data, meta = live_bch
    for name, m in meta.items():
        ext = gb.GF(2 ** m["m"], irreducible_poly=m["irreducible_poly"], primitive_element=m["primitive_element"])
        if m["by_k"]:
            bch = gb.BCH(m["n"], m["k"], extension_field=ext, alpha=m["alpha"], c=m["c"])
        else:
            bch = gb.BCH(m["n"], d=m["d"], extension_field=ext, alpha=m["alpha"], c=m["c"])
        assert (bch.n, bch.k, bch.d, bch.t) == (m["n"], m["k"], m["d"], m["t"]), name
        assert np.array_equal(np.asarray(bch.generator_poly), data[f"{name}/generator_poly"]), name
        assert np.array_equal(np.asarray(bch.roots), data[f"{name}/roots"]), name
        assert bch.field.order == 2 and bch.extension_field is ext
    assert gb.BCH(15, 7).d == 5 and gb.BCH(15, d=5).k == 7  # default extension field GF(2^4), x^4 + x + 1
    with pytest.raises(ValueError):
        gb.BCH(15, 8)
    with pytest.raises(ValueError):
        gb.BCH(15, 7, 7)
    with pytest.raises(ValueError):
        gb.BCH(15)
    with pytest.raises(NotImplementedError):
        gb.BCH(13, 4, field=gb.GF(3))


def test_poly_holder_and_argument_checks_need_no_gpu():
    """Poly construction (sparse storage, negative literals, asc/desc order) and the argument errors raised before any
    kernel is reached (roots_jit.__call__ _polys/_dense.py:448-452, Poly.Degrees _polys/_poly.py:426-520)."""
    GF = gb.GF(7)
    p = gb.Poly([1, 0, -5, 0], field=GF)
    assert p.degree == 3 and p.nonzero_degrees.tolist() == [3, 1] and np.asarray(p.nonzero_coeffs).tolist() == [1, 2]
    assert np.asarray(p.coeffs).tolist() == [1, 0, 2, 0]
    assert gb.Poly([0, 2, 0, 1], field=GF, order="asc") == p
    q = gb.Poly.Degrees([1, 9, 5], [1, 1, -1], field=GF)
    assert q.nonzero_degrees.tolist() == [9, 5, 1] and np.asarray(q.nonzero_coeffs).tolist() == [1, 6, 1]
    assert gb.Poly([1, 1]).field is gb.GF(2)  # default field, like galois.Poly
    G8 = gb.GF(2**8)
    assert np.asarray(gb.Poly([3, -7], field=G8).nonzero_coeffs).tolist() == [3, 7]  # -c = c in characteristic 2
    with pytest.raises(ValueError):
        gb.Poly.Degrees([3, 3], [1, 1], field=GF)
    with pytest.raises(ValueError):
        gb.Poly.Degrees([-1], [1], field=GF)
    with pytest.raises(TypeError):
        gb.poly_roots([2, 0], GF([1, 1]))
    with pytest.raises(TypeError):
        gb.poly_roots(np.array([2, 0]), np.array([1, 1]))
    with pytest.raises(ValueError):
        gb.poly_roots(np.array([], dtype=np.int64), GF([]))
    with pytest.raises(TypeError):
        p.roots(multiplicity=1)
    with pytest.raises(NotImplementedError):
        gb.poly_roots(np.array([1, 0]), gb.GF(2**100)([1, 1]))


def test_linalg_argument_checks_need_no_gpu():
    """Shape / type errors of the elimination family mirror the reference's and are raised before any kernel runs
    (tests/fields/test_linalg.py:291-299, 321-351, 395-403, 415-423, 435-457, 471-539)."""
    GF = gb.GF(2**8)
    v, t = GF.Random(5, seed=1), GF.Random((2, 2, 2), seed=2)
    for bad in (v, t):
        with pytest.raises(ValueError):
            bad.row_reduce()
        with pytest.raises(ValueError):
            bad.lu_decompose()
        with pytest.raises(ValueError):
            bad.plu_decompose()
        for method in ("row_space", "column_space", "left_null_space", "null_space"):
            with pytest.raises(ValueError):
                getattr(bad, method)()
    A = GF.Random((2, 3), seed=3)
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.inv(A)
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.det(A)
    S = GF.Random((3, 3), seed=4)
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.solve(A, GF.Random(3, seed=5))
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.solve(S, GF.Random((3, 2, 2), seed=5))
    with pytest.raises(np.linalg.LinAlgError):
        np.linalg.solve(S, GF.Random(4, seed=5))
    with pytest.raises(TypeError):
        np.linalg.solve(S, np.asarray(v[:3]))
    with pytest.raises(ValueError):
        S.row_reduce(eye="middle")
    with pytest.raises(TypeError):
        np.dot(S, np.asarray(S))
    with pytest.raises(ValueError):
        np.inner(GF.Random((2, 3), seed=6), GF.Random((2, 4), seed=7))


def test_fft_size_checks_need_no_gpu():
    """Any n dividing order - 1 is a legal transform size now (mixed radix); everything else is a ValueError."""
    with pytest.raises(ValueError):
        np.fft.fft(gb.GF(31).Random(7, seed=1))
    with pytest.raises(ValueError):
        np.fft.fft(gb.GF(2**8).Random(16, seed=1))
    with pytest.raises(NotImplementedError):
        np.fft.fft(gb.GF(2**100).Random(3, seed=1))


def test_getitem_setitem_astype_stay_in_the_field():
    """scalar indexing, assignment and casts follow _domains/_array.py:413-451 (no GPU needed)."""
    import galois_b200 as gb

    GF = gb.GF(7)
    x = GF([1, 2, 3, 4])
    assert type(x[0]) is GF and x[0].ndim == 0 and x[0].dtype == x.dtype
    assert type(x[1:3]) is GF
    x[0] = 6
    x[1:3] = [0, 5]
    x[3] = GF(2)
    assert np.array_equal(np.asarray(x), [6, 0, 5, 2])
    with pytest.raises(ValueError):
        x[0] = 100
    with pytest.raises(ValueError):
        x[0:2] = [1, 7]
    with pytest.raises(ValueError):
        x[0] = -1
    with pytest.raises(TypeError):
        x[0] = 1.5
    with pytest.raises(TypeError):
        x[0] = gb.GF(5)(1)
    assert np.array_equal(np.asarray(x), [6, 0, 5, 2])  # failed assignments left the array untouched
    assert x.astype(np.int64).dtype == np.int64 and type(x.astype(np.int64)) is GF
    with pytest.raises(TypeError):
        x.astype(np.float64)
    with pytest.raises(TypeError):
        gb.GF(2**31 - 1)([1, 2]).astype(np.uint8)
    W = gb.GF(2**100)
    w = W([1 << 90, 5])
    assert type(w[0]) is W and int(w[0]) == 1 << 90


def test_reduce_exponent_keeps_the_sign():
    """a huge negative exponent must stay negative so that 0 ** e still raises in the kernel."""
    import galois_b200 as gb

    GF = gb.GF(31)
    assert GF._reduce_exponent(-(2**70)) < 0
    assert GF._reduce_exponent(-(2**70)) % 30 == (-(2**70)) % 30
    assert GF._reduce_exponent(2**70) % 30 == (2**70) % 30 and GF._reduce_exponent(2**70) > 0
    assert GF._reduce_exponent(-30 * 2**70) == -30  # multiple of order - 1: a ** e = 1 for a != 0, still negative
    assert GF._reduce_exponent(5) == 5
