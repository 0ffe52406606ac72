"""
Generate the golden fixtures under tests/golden/ from the REAL reference (mhostetter/galois,
mounted read-only at /root/reference).  Run in the build container only — the GPU box has no
/root/reference; it only sees the committed .npz files.

    NUMBA_CACHE_DIR=/tmp/numba python tests/golden/make_golden.py

Two families of fixtures:
  ref_*.npz   conversions of the reference's own third-party vectors (Sage / SymPy pickles under
              /root/reference/tests/**/data), i.e. what the reference's test-suite pins.
  live_*.npz  inputs + outputs of the live reference (its Numba/NumPy path) on seeded inputs for
              everything those vectors do not cover: GF(998244353) NTTs, RS(255,223), GF(2^128), ...
"""

from __future__ import annotations

import hashlib
import json
import os
import pickle
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
import warnings  # noqa: E402

warnings.filterwarnings("ignore")
import galois  # noqa: E402

OUT = Path(__file__).resolve().parent
MASK64 = (1 << 64) - 1


def to_limbs(arr) -> np.ndarray:
    flat = [int(v) for v in np.asarray(arr, dtype=object).ravel()]
    out = np.empty((len(flat), 2), dtype=np.uint64)
    for i, v in enumerate(flat):
        out[i, 0] = v & MASK64
        out[i, 1] = v >> 64
    return out.reshape(np.asarray(arr, dtype=object).shape + (2,))


def poly_int(coeffs_desc, p) -> int:
    v = 0
    for c in coeffs_desc:
        v = v * p + int(c)
    return v


# --------------------------------------------------------------------------- reference's own vectors
def convert_field_vectors():
    fields = ["GF(2)", "GF(2^2)", "GF(2^3)", "GF(2^8)", "GF(2^32)", "GF(2^100)", "GF(5)", "GF(7)", "GF(31)",
              "GF(3191)", "GF(2147483647)"]
    ops = ["add", "subtract", "multiply", "divide", "additive_inverse", "multiplicative_inverse", "power",
           "convolve", "matrix_multiply"]
    out = {}
    meta = {}
    for fld in fields:
        folder = REF / "tests" / "fields" / "data" / fld
        props = json.load(open(folder / "properties.json"))
        p = props["characteristic"]
        meta[fld] = {
            "characteristic": p,
            "degree": props["degree"],
            "order": props["order"],
            "primitive_element": props["primitive_element"] if isinstance(props["primitive_element"], int)
            else poly_int(props["primitive_element"], p),
            "irreducible_poly": poly_int(props["irreducible_poly"], p),
        }
        wide = props["order"] > 2**63
        for op in ops:
            d = pickle.load(open(folder / f"{op}.pkl", "rb"))
            if op == "power" and wide:
                # keep the exponents an int64 kernel argument can carry (the C ABI takes int64 exponents)
                keep = [j for j, v in enumerate(d["Y"]) if -(2**63) < int(v) < 2**63]
                d = {"X": d["X"], "Y": np.array([int(d["Y"][j]) for j in keep], dtype=np.int64),
                     "Z": np.asarray(d["Z"], dtype=object)[:, keep]}
            for key in ("X", "Y", "Z"):
                if key not in d:
                    continue
                val = d[key]
                if isinstance(val, list):
                    for i, item in enumerate(val):
                        a = np.asarray(item)
                        out[f"{fld}/{op}/{key}{i}"] = to_limbs(a) if wide and not (op == "power" and key == "Y") else a.astype(np.int64)
                    out[f"{fld}/{op}/count"] = np.array(len(val))
                else:
                    a = np.asarray(val)
                    if wide and not (op == "power" and key == "Y"):
                        out[f"{fld}/{op}/{key}"] = to_limbs(a)
                    else:
                        out[f"{fld}/{op}/{key}"] = a.astype(np.int64)
    np.savez_compressed(OUT / "ref_fields.npz", **out)
    json.dump(meta, open(OUT / "ref_fields_meta.json", "w"), indent=1)
    print("ref_fields.npz", len(out), "arrays")


def convert_rs_vectors():
    out = {}
    names = []
    folder = REF / "tests" / "codes" / "data" / "reed_solomon"
    for f in sorted(folder.glob("n15_*_sys.pkl")):
        d = pickle.load(open(f, "rb"))
        name = f.stem
        names.append(name)
        out[f"{name}/params"] = np.array([d["q"], d["n"], d["k"], d["d"], d["alpha"], d["c"]], dtype=np.int64)
        out[f"{name}/G"] = np.asarray(d["G"], dtype=np.int64)
        out[f"{name}/H"] = np.asarray(d["H"], dtype=np.int64)
        out[f"{name}/messages"] = np.asarray(d["encode"]["messages"], dtype=np.int64)
        out[f"{name}/codewords"] = np.asarray(d["encode"]["codewords"], dtype=np.int64)
        out[f"{name}/short_messages"] = np.asarray(d["encode_shortened"]["messages"], dtype=np.int64)
        out[f"{name}/short_codewords"] = np.asarray(d["encode_shortened"]["codewords"], dtype=np.int64)
    out["names"] = np.array(names)
    np.savez_compressed(OUT / "ref_rs.npz", **out)
    print("ref_rs.npz", len(names), "codes")


# --------------------------------------------------------------------------- live reference outputs
def live_elementwise():
    out = {}
    rng = np.random.default_rng(20260923)

    def grid(GF, name, n, dtypes):
        order = GF.order
        edge = np.array([0, 1, 2, order - 2, order - 1], dtype=np.int64) % order
        for dt in dtypes:
            x = np.concatenate([edge.repeat(5), rng.integers(0, order, n)]).astype(dt)
            y = np.concatenate([np.tile(edge, 5), rng.integers(0, order, n)]).astype(dt)
            ynz = y.copy()
            ynz[ynz == 0] = 1
            X, Y, YNZ = GF(x), GF(y), GF(ynz)
            e = np.concatenate([np.array([0, 1, -1, 2, -2, order - 1, -(order - 1), order, 2**40, -(2**40) + 1]),
                                rng.integers(-1000, 1000, x.size - 10)]).astype(np.int64)
            xe = x.copy()
            xe[(xe == 0) & (e < 0)] = 1
            key = f"{name}/{np.dtype(dt).name}"
            out[f"{key}/x"], out[f"{key}/y"], out[f"{key}/ynz"] = x, y, ynz
            out[f"{key}/add"] = np.asarray(X + Y)
            out[f"{key}/sub"] = np.asarray(X - Y)
            out[f"{key}/mul"] = np.asarray(X * Y)
            out[f"{key}/div"] = np.asarray(X / YNZ)
            out[f"{key}/neg"] = np.asarray(-X)
            out[f"{key}/inv"] = np.asarray(np.reciprocal(YNZ))
            out[f"{key}/square"] = np.asarray(X**2)
            out[f"{key}/e"], out[f"{key}/xe"] = e, xe
            out[f"{key}/pow"] = np.asarray(GF(xe) ** e)
            out[f"{key}/pow3"] = np.asarray(X**3)
            out[f"{key}/powm1"] = np.asarray(YNZ**-1)
            out[f"{key}/mul_scalar7"] = np.asarray(X * GF(7 % order))
            out[f"{key}/scalar5_sub"] = np.asarray(GF(5 % order) - X)
            out[f"{key}/scalar3_div"] = np.asarray(GF(3 % order) / YNZ)

    grid(galois.GF(2**8), "GF(2^8)", 4000, [np.uint8, np.int16, np.int64])
    grid(galois.GF(2**31 - 1), "GF(2147483647)", 4001, [np.uint32, np.int32, np.int64])
    grid(galois.GF(998244353), "GF(998244353)", 2003, [np.uint32, np.int64])
    grid(galois.GF(2**12), "GF(2^12)", 3001, [np.uint16, np.int32])
    grid(galois.GF(2**16), "GF(2^16)", 2001, [np.uint16, np.int64])
    grid(galois.GF(2**32), "GF(2^32)", 1501, [np.uint32, np.int64])
    grid(galois.GF(251), "GF(251)", 1001, [np.uint8])
    grid(galois.GF(65537), "GF(65537)", 1001, [np.uint32])
    grid(galois.GF(2), "GF(2)", 301, [np.uint8])
    grid(galois.GF(3), "GF(3)", 301, [np.uint8])
    # fields whose constants the fixtures must carry
    meta = {}
    for order in [2**8, 2**12, 2**16, 2**32]:
        GF = galois.GF(order)
        meta[f"GF(2^{GF.degree})"] = {"irreducible_poly": int(GF.irreducible_poly),
                                      "primitive_element": int(GF.primitive_element)}
    json.dump(meta, open(OUT / "live_elementwise_meta.json", "w"), indent=1)
    np.savez_compressed(OUT / "live_elementwise.npz", **out)
    print("live_elementwise.npz", len(out), "arrays")


def live_gf2_128():
    out = {}
    GF = galois.GF(2**128, irreducible_poly="x^128 + x^7 + x^2 + x + 1", primitive_element=2, verify=False)
    rng = np.random.default_rng(5)
    n = 600
    lo, hi = rng.integers(0, 2**64, (2, n), dtype=np.uint64)
    lo2, hi2 = rng.integers(0, 2**64, (2, n), dtype=np.uint64)
    # edge elements first
    edges = [0, 1, 2, (1 << 128) - 1, 1 << 127, (1 << 64) - 1, 1 << 64, 0x87]
    a_int = edges + [int(l) | (int(h) << 64) for l, h in zip(lo, hi)]
    b_int = list(reversed(edges)) + [int(l) | (int(h) << 64) for l, h in zip(lo2, hi2)]
    A = GF(np.array(a_int, dtype=object))
    B = GF(np.array(b_int, dtype=object))
    out["a"], out["b"] = to_limbs(a_int), to_limbs(b_int)
    out["mul"] = to_limbs(np.asarray(A * B, dtype=object))
    out["add"] = to_limbs(np.asarray(A + B, dtype=object))
    out["square"] = to_limbs(np.asarray(A**2, dtype=object))
    ni = 48
    bnz = [v if v else 1 for v in b_int[:ni]]
    BN = GF(np.array(bnz, dtype=object))
    out["inv_in"] = to_limbs(bnz)
    out["inv"] = to_limbs(np.asarray(np.reciprocal(BN), dtype=object))
    out["div"] = to_limbs(np.asarray(A[:ni] / BN, dtype=object))
    e = np.array([0, 1, 2, 3, -1, -2, 5, 17, 255, 1000, -3, 2**40] + list(rng.integers(-50, 50, ni - 12)), dtype=np.int64)
    pe = [v if v else 1 for v in a_int[:ni]]
    out["pow_in"], out["pow_e"] = to_limbs(pe), e
    out["pow"] = to_limbs(np.asarray(GF(np.array(pe, dtype=object)) ** e, dtype=object))
    # a second wide field: the reference's own GF(2^100)
    GF100 = galois.GF(2**100)
    a100 = [int(v) for v in rng.integers(0, 2**50, 200)]
    a100 = [(x << 50) | y for x, y in zip(a100, reversed(a100))]
    b100 = [(x * 0x9E3779B97F4A7C15 + 12345) % (1 << 100) for x in a100]
    A100, B100 = GF100(np.array(a100, dtype=object)), GF100(np.array(b100, dtype=object))
    out["gf100_poly"] = to_limbs([int(GF100.irreducible_poly)])
    out["gf100_a"], out["gf100_b"] = to_limbs(a100), to_limbs(b100)
    out["gf100_mul"] = to_limbs(np.asarray(A100 * B100, dtype=object))
    b100nz = [v if v else 1 for v in b100[:24]]
    out["gf100_inv_in"] = to_limbs(b100nz)
    out["gf100_inv"] = to_limbs(np.asarray(np.reciprocal(GF100(np.array(b100nz, dtype=object))), dtype=object))
    np.savez_compressed(OUT / "live_gf2_128.npz", **out)
    print("live_gf2_128.npz")


def live_ntt():
    out = {}
    meta = {}
    rng = np.random.default_rng(3)
    cases = [(998244353, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 2**14, 2**16]),
             (12289, [2, 8, 1024, 4096]), (7340033, [64, 2048, 2**15]), (469762049, [16, 2**13]),
             (2013265921, [4, 512, 2**12, 2**17]), (257, [256]), (65537, [2**16]), (17, [16]), (5, [4])]
    for p, sizes in cases:
        GF = galois.GF(p)
        for n in sizes:
            x = GF.Random(n, seed=int(rng.integers(1 << 30)))
            X = galois.ntt(x)
            xi = galois.intt(X)
            assert np.array_equal(xi, x)
            Xu = galois.intt(x, scaled=False)
            key = f"p{p}/n{n}"
            out[f"{key}/x"] = np.asarray(x).astype(np.int64)
            out[f"{key}/ntt"] = np.asarray(X).astype(np.int64)
            out[f"{key}/intt"] = np.asarray(galois.intt(x)).astype(np.int64)
            out[f"{key}/intt_unscaled"] = np.asarray(Xu).astype(np.int64)
            meta[key] = {"omega": int(GF.primitive_root_of_unity(n)), "primitive_element": int(GF.primitive_element)}
    # zero padded / truncated sizes (galois.ntt(x, size=...))
    GF = galois.GF(998244353)
    x = GF.Random(100, seed=11)
    out["pad/x"] = np.asarray(x).astype(np.int64)
    out["pad/ntt128"] = np.asarray(galois.ntt(x, size=128)).astype(np.int64)
    # the headline size: N = 2^20.  Inputs are regenerated from the seed; outputs pinned by digest + samples.
    for seed in (3, 4):
        x = GF.Random(2**20, seed=seed)
        X = np.asarray(galois.ntt(x)).astype(np.uint32)
        Xi = np.asarray(galois.intt(x)).astype(np.uint32)
        meta[f"p998244353/n1048576/seed{seed}"] = {
            "omega": int(GF.primitive_root_of_unity(2**20)),
            "x_sha256": hashlib.sha256(np.asarray(x).astype(np.uint32).tobytes()).hexdigest(),
            "ntt_sha256": hashlib.sha256(X.tobytes()).hexdigest(),
            "intt_sha256": hashlib.sha256(Xi.tobytes()).hexdigest(),
        }
        out[f"big/seed{seed}/ntt_head"] = X[:64].astype(np.int64)
        out[f"big/seed{seed}/ntt_stride"] = X[:: 2**12].astype(np.int64)
        out[f"big/seed{seed}/intt_stride"] = Xi[:: 2**12].astype(np.int64)
    json.dump(meta, open(OUT / "live_ntt_meta.json", "w"), indent=1)
    np.savez_compressed(OUT / "live_ntt.npz", **out)
    print("live_ntt.npz", len(out), "arrays")


def live_linalg():
    out = {}
    rng = np.random.default_rng(8)
    for name, GF in [("GF(998244353)", galois.GF(998244353)), ("GF(7)", galois.GF(7)), ("GF(2^8)", galois.GF(2**8)),
                     ("GF(2147483647)", galois.GF(2**31 - 1)), ("GF(2^32)", galois.GF(2**32))]:
        for (na, nb) in [(1, 1), (5, 3), (33, 33), (300, 200), (64, 1000)]:
            a = GF.Random(na, seed=int(rng.integers(1 << 30)))
            b = GF.Random(nb, seed=int(rng.integers(1 << 30)))
            out[f"conv/{name}/{na}x{nb}/a"] = np.asarray(a).astype(np.int64)
            out[f"conv/{name}/{na}x{nb}/b"] = np.asarray(b).astype(np.int64)
            out[f"conv/{name}/{na}x{nb}/c"] = np.asarray(np.convolve(a, b)).astype(np.int64)
        for (M, K, N) in [(1, 1, 1), (7, 5, 3), (20, 30, 10), (64, 223, 32)]:
            A = GF.Random((M, K), seed=int(rng.integers(1 << 30)))
            B = GF.Random((K, N), seed=int(rng.integers(1 << 30)))
            out[f"matmul/{name}/{M}x{K}x{N}/A"] = np.asarray(A).astype(np.int64)
            out[f"matmul/{name}/{M}x{K}x{N}/B"] = np.asarray(B).astype(np.int64)
            out[f"matmul/{name}/{M}x{K}x{N}/C"] = np.asarray(A @ B).astype(np.int64)
        A = GF.Random((3, 4, 6), seed=1)
        B = GF.Random((3, 6, 5), seed=2)
        out[f"matmul/{name}/batched/A"], out[f"matmul/{name}/batched/B"] = np.asarray(A).astype(np.int64), np.asarray(B).astype(np.int64)
        out[f"matmul/{name}/batched/C"] = np.asarray(A @ B).astype(np.int64)
    # Berlekamp-Massey + polynomial evaluation over GF(2^8)
    GF = galois.GF(2**8)
    for i in range(12):
        s = GF.Random(int(rng.integers(4, 40)), seed=100 + i)
        c = galois.berlekamp_massey(s)
        out[f"bm/{i}/s"] = np.asarray(s).astype(np.int64)
        out[f"bm/{i}/c"] = np.asarray(c.coeffs).astype(np.int64)
        poly = galois.Poly(GF.Random(int(rng.integers(1, 30)), seed=200 + i))
        pts = GF.Random(50, seed=300 + i)
        out[f"eval/{i}/coeffs"] = np.asarray(poly.coeffs).astype(np.int64)
        out[f"eval/{i}/x"] = np.asarray(pts).astype(np.int64)
        out[f"eval/{i}/y"] = np.asarray(poly(pts)).astype(np.int64)
    np.savez_compressed(OUT / "live_linalg.npz", **out)
    print("live_linalg.npz", len(out), "arrays")


def corrupt(C, rng, n_err, n_era):
    """Returns (received, erasure mask): n_err symbol errors and n_era erasures per row, disjoint.
    Fields are GF(2^m): an error adds (xors) a non-zero symbol; an erased symbol is replaced by noise."""
    N, n = C.shape
    order = int(type(C).order)
    R = np.asarray(C).copy()
    E = np.zeros((N, n), dtype=bool)
    for i in range(N):
        pos = rng.choice(n, n_err + n_era, replace=False)
        for ppos in pos[:n_err]:
            R[i, ppos] ^= rng.integers(1, order)
        for ppos in pos[n_err:]:
            E[i, ppos] = True
            R[i, ppos] = rng.integers(0, order)
    return R, E


def live_rs():
    out = {}
    meta = {}
    rng = np.random.default_rng(7)
    codes = [
        ("rs255_223", dict(n=255, k=223), 48),
        ("rs255_239", dict(n=255, k=239), 16),
        ("rs255_191_c0", dict(n=255, k=191, c=0), 8),
        ("rs63_55", dict(n=63, k=55), 16),
        ("rs15_9", dict(n=15, k=9), 32),
        ("rs31_25_alpha", dict(n=31, k=25, alpha=None), 16),
        ("rs85_65", dict(n=85, k=65, field=galois.GF(2**8)), 8),
    ]
    for name, kw, N in codes:
        kw = dict(kw)
        if kw.get("alpha", 0) is None:
            GF = galois.GF(2**5)
            kw["field"] = GF
            kw["alpha"] = GF.primitive_elements[3]
        rs = galois.ReedSolomon(**kw)
        GF = rs.field
        t = (rs.d - 1) // 2
        meta[name] = {"n": rs.n, "k": rs.k, "d": rs.d, "c": int(rs.c), "alpha": int(rs.alpha), "t": t,
                      "m": GF.degree, "irreducible_poly": int(GF.irreducible_poly),
                      "primitive_element": int(GF.primitive_element)}
        out[f"{name}/generator_poly"] = np.asarray(rs.generator_poly.coeffs).astype(np.int64)
        out[f"{name}/roots"] = np.asarray(rs.roots).astype(np.int64)
        out[f"{name}/G"] = np.asarray(rs.G).astype(np.int64)
        M = GF.Random((N, rs.k), seed=int(rng.integers(1 << 30)))
        C = rs.encode(M)
        out[f"{name}/messages"] = np.asarray(M).astype(np.int64)
        out[f"{name}/codewords"] = np.asarray(C).astype(np.int64)
        patterns = {
            "clean": (0, 0), "t_errors": (t, 0), "one_error": (1, 0), "max_erasures": (0, rs.d - 1),
            "mixed": (t // 2, rs.d - 1 - 2 * (t // 2)), "mixed_small": (1, 2), "too_many": (rs.d - 1, 0),
            "t_plus_1": (t + 1, 0),
        }
        for pname, (ne, nu) in patterns.items():
            if ne + nu > rs.n:
                continue
            R, E = corrupt(C, rng, ne, nu)
            R = GF(R)
            dec, nerr = rs.decode(R, erasures=E if nu else None, output="codeword", errors=True)
            out[f"{name}/{pname}/received"] = np.asarray(R).astype(np.int64)
            out[f"{name}/{pname}/erasures"] = E
            out[f"{name}/{pname}/decoded"] = np.asarray(dec).astype(np.int64)
            out[f"{name}/{pname}/n_errors"] = np.asarray(nerr).astype(np.int64)
        # shortened code: ks message symbols
        s = max(1, rs.k // 3)
        Ms = GF.Random((N, rs.k - s), seed=int(rng.integers(1 << 30)))
        Cs = rs.encode(Ms)
        out[f"{name}/short/messages"] = np.asarray(Ms).astype(np.int64)
        out[f"{name}/short/codewords"] = np.asarray(Cs).astype(np.int64)
        for pname, (ne, nu) in {"t_errors": (t, 0), "mixed": (t // 2, 2), "t_plus_1": (t + 1, 0)}.items():
            R, E = corrupt(Cs, rng, ne, nu)
            R = GF(R)
            dec, nerr = rs.decode(R, erasures=E if nu else None, output="codeword", errors=True)
            out[f"{name}/short/{pname}/received"] = np.asarray(R).astype(np.int64)
            out[f"{name}/short/{pname}/erasures"] = E
            out[f"{name}/short/{pname}/decoded"] = np.asarray(dec).astype(np.int64)
            out[f"{name}/short/{pname}/n_errors"] = np.asarray(nerr).astype(np.int64)
    json.dump(meta, open(OUT / "live_rs_meta.json", "w"), indent=1)
    np.savez_compressed(OUT / "live_rs.npz", **out)
    print("live_rs.npz", len(out), "arrays")


def convert_bch_vectors():
    """The reference's Sage vectors for binary BCH codes (tests/codes/data/bch/n15_*_sys.pkl with q = 2)."""
    out, names = {}, []
    folder = REF / "tests" / "codes" / "data" / "bch"
    for f in sorted(folder.glob("*_sys.pkl")):
        d = pickle.load(open(f, "rb"))
        if d["q"] != 2:
            continue
        name = f.stem
        names.append(name)
        out[f"{name}/params"] = np.array([d["q"], d["n"], d["k"], d["d"], d["alpha"], d["c"]], dtype=np.int64)
        out[f"{name}/G"] = np.asarray(d["G"], dtype=np.int64)
        out[f"{name}/messages"] = np.asarray(d["encode"]["messages"], dtype=np.int64)
        out[f"{name}/codewords"] = np.asarray(d["encode"]["codewords"], dtype=np.int64)
        if d.get("encode_shortened") and "messages" in d["encode_shortened"]:
            out[f"{name}/short_messages"] = np.asarray(d["encode_shortened"]["messages"], dtype=np.int64)
            out[f"{name}/short_codewords"] = np.asarray(d["encode_shortened"]["codewords"], dtype=np.int64)
    out["names"] = np.array(names)
    np.savez_compressed(OUT / "ref_bch.npz", **out)
    print("ref_bch.npz", len(names), "codes")


def live_bch():
    """Binary BCH codes on the live reference: construction (by k and by d), encode, decode under error / erasure
    patterns including uncorrectable ones, shortened codes."""
    out, meta = {}, {}
    rng = np.random.default_rng(17)
    codes = [("bch15_7", dict(n=15, k=7), 32), ("bch15_5", dict(n=15, k=5), 16), ("bch31_16", dict(n=31, k=16), 16),
             ("bch63_36", dict(n=63, k=36), 16), ("bch127_64", dict(n=127, k=64), 8),
             ("bch255_131", dict(n=255, k=131), 24), ("bch255_d37", dict(n=255, d=37), 8),
             ("bch255_223_c3", dict(n=255, d=9, c=3), 8), ("bch51_27", dict(n=51, d=9, extension_field=galois.GF(2**8)), 8)]
    GF2 = galois.GF(2)

    def store(key, bch, R, E):
        """Decode row by row: beyond the design distance the reference can miscorrect to a word with non-binary
        symbols and then raises ValueError from `.view(GF(2))` (_bch.py:1300, _fields/_array.py:177); record that."""
        Rn = np.asarray(R)
        dec = Rn.astype(np.int64).copy()
        nerr = np.zeros(Rn.shape[0], dtype=np.int64)
        raised = np.zeros(Rn.shape[0], dtype=bool)
        for i in range(Rn.shape[0]):
            try:
                d_i, n_i = bch.decode(R[i], erasures=E[i] if E is not None else None, output="codeword", errors=True)
                dec[i], nerr[i] = np.asarray(d_i).astype(np.int64), n_i
            except ValueError:
                raised[i] = True
        out[f"{key}/received"] = Rn.astype(np.int64)
        out[f"{key}/erasures"] = E if E is not None else np.zeros(Rn.shape, dtype=bool)
        out[f"{key}/decoded"] = dec
        out[f"{key}/n_errors"] = nerr
        out[f"{key}/raised"] = raised

    for name, kw, N in codes:
        bch = galois.BCH(**kw)
        EF = bch.extension_field
        t = (bch.d - 1) // 2
        meta[name] = {"n": bch.n, "k": bch.k, "d": bch.d, "c": int(bch.c), "alpha": int(bch.alpha), "t": t,
                      "m": EF.degree, "irreducible_poly": int(EF.irreducible_poly),
                      "primitive_element": int(EF.primitive_element), "by_k": "k" in kw}
        out[f"{name}/generator_poly"] = np.asarray(bch.generator_poly.coeffs).astype(np.int64)
        out[f"{name}/roots"] = np.asarray(bch.roots).astype(np.int64)
        out[f"{name}/G"] = np.asarray(bch.G).astype(np.int64)
        M = GF2.Random((N, bch.k), seed=int(rng.integers(1 << 30)))
        C = bch.encode(M)
        out[f"{name}/messages"] = np.asarray(M).astype(np.int64)
        out[f"{name}/codewords"] = np.asarray(C).astype(np.int64)

        def corrupt2(C, n_err, n_era):
            R = np.asarray(C).copy()
            E = np.zeros(R.shape, dtype=bool)
            for i in range(R.shape[0]):
                pos = rng.choice(R.shape[1], n_err + n_era, replace=False)
                R[i, pos[:n_err]] ^= 1
                E[i, pos[n_err:]] = True
                R[i, pos[n_err:]] = rng.integers(0, 2, n_era)
            return GF2(R), E

        patterns = {"clean": (0, 0), "t_errors": (t, 0), "one_error": (min(1, t), 0), "max_erasures": (0, bch.d - 1),
                    "mixed": (t // 2, bch.d - 1 - 2 * (t // 2)), "t_plus_1": (t + 1, 0), "too_many": (min(bch.n, bch.d + 3), 0),
                    "era_plus_1": (1, bch.d - 1)}
        for pname, (ne, nu) in patterns.items():
            if ne + nu > bch.n:
                continue
            R, E = corrupt2(C, ne, nu)
            store(f"{name}/{pname}", bch, R, E if nu else None)
        s = max(1, bch.k // 3)
        Ms = GF2.Random((N, bch.k - s), seed=int(rng.integers(1 << 30)))
        Cs = bch.encode(Ms)
        out[f"{name}/short/messages"] = np.asarray(Ms).astype(np.int64)
        out[f"{name}/short/codewords"] = np.asarray(Cs).astype(np.int64)
        for pname, (ne, nu) in {"t_errors": (t, 0), "t_plus_1": (t + 1, 0)}.items():
            R, E = corrupt2(Cs, ne, nu)
            store(f"{name}/short/{pname}", bch, R, None)
    json.dump(meta, open(OUT / "live_bch_meta.json", "w"), indent=1)
    np.savez_compressed(OUT / "live_bch.npz", **out)
    print("live_bch.npz", len(out), "arrays")



# --------------------------------------------------------------------------- §8(f) rows: roots, mixed-radix FFT, dense linalg
LINALG_FIELDS = ["GF(2)", "GF(2^2)", "GF(2^3)", "GF(2^8)", "GF(2^32)", "GF(5)", "GF(7)", "GF(31)", "GF(3191)",
                 "GF(2147483647)"]


def convert_linalg_vectors():
    """Sage vectors for row_reduce / lu / plu / inv / det / solve / row, column and null spaces (tests/fields/data)."""
    ops = {"row_reduce": ("X", "Z"), "lu_decompose": ("X", "L", "U"), "plu_decompose": ("X", "P", "L", "U"),
           "matrix_inverse": ("X", "Z"), "matrix_determinant": ("X", "Z"), "matrix_solve": ("X", "Y", "Z"),
           "row_space": ("X", "Z"), "column_space": ("X", "Z"), "left_null_space": ("X", "Z"), "null_space": ("X", "Z")}
    out = {}
    for fld in LINALG_FIELDS:
        folder = REF / "tests" / "fields" / "data" / fld
        for op, keys in ops.items():
            d = pickle.load(open(folder / f"{op}.pkl", "rb"))
            out[f"{fld}/{op}/count"] = np.array(len(d[keys[0]]))
            for key in keys:
                for i, item in enumerate(d[key]):
                    out[f"{fld}/{op}/{key}{i}"] = np.asarray(item).astype(np.int64)
    np.savez_compressed(OUT / "ref_linalg.npz", **out)
    print("ref_linalg.npz", len(out), "arrays")


def convert_roots_vectors():
    """Sage vectors for Poly.roots (tests/polys/data/*/roots.pkl): degree-descending coefficients -> sorted roots."""
    out = {}
    for fld in LINALG_FIELDS:
        d = pickle.load(open(REF / "tests" / "polys" / "data" / fld / "roots.pkl", "rb"))
        out[f"{fld}/count"] = np.array(len(d["X"]))
        for i, (x, r, m) in enumerate(zip(d["X"], d["R"], d["M"])):
            out[f"{fld}/X{i}"] = np.asarray(x, dtype=np.int64)
            out[f"{fld}/R{i}"] = np.asarray(r, dtype=np.int64)
            out[f"{fld}/M{i}"] = np.asarray(m, dtype=np.int64)
    np.savez_compressed(OUT / "ref_roots.npz", **out)
    print("ref_roots.npz", len(out), "arrays")


def live_next_rows():
    """Live-reference outputs for the §8(f) rows at sizes the Sage vectors do not reach."""
    out = {}
    meta = {}
    rng = np.random.default_rng(2024)

    # ---- Poly.roots: dense and sparse polynomials, roots with multiplicity, the x^q - x extreme, no roots
    def roots_case(key, GF, poly):
        out[f"roots/{key}/degrees"] = np.asarray(poly.nonzero_degrees, dtype=np.int64)
        out[f"roots/{key}/coeffs"] = np.asarray(poly.nonzero_coeffs, dtype=np.int64)
        out[f"roots/{key}/roots"] = np.asarray(poly.roots(), dtype=np.int64)

    fields = {"gf256": galois.GF(2**8), "gf65536": galois.GF(2**16), "gf7919": galois.GF(7919),
              "gf998244353_small": None, "gf2": galois.GF(2), "gf31": galois.GF(31)}
    meta["roots"] = {}
    for name, GF in fields.items():
        if GF is None:
            continue
        meta["roots"][name] = {"order": int(GF.order), "irreducible_poly": int(GF.irreducible_poly),
                               "primitive_element": int(GF.primitive_element)}
        q = GF.order
        r = GF.Random(min(6, q - 1), seed=int(rng.integers(1 << 30)), low=1)
        roots_case(f"{name}/from_roots", GF, galois.Poly.Roots(r, multiplicities=[1 + i % 3 for i in range(r.size)]))
        roots_case(f"{name}/random_dense", GF, galois.Poly.Random(17, seed=int(rng.integers(1 << 30)), field=GF))
        roots_case(f"{name}/sparse", GF, galois.Poly.Degrees([q + 7, 5, 1], [1, int(GF(q - 1)), 1], field=GF))
        roots_case(f"{name}/x_q_minus_x", GF, galois.Poly.Degrees([q, 1], [1, int(-GF(1))], field=GF))
        roots_case(f"{name}/x_pow", GF, galois.Poly.Degrees([9], [1], field=GF))
        roots_case(f"{name}/constant", GF, galois.Poly([int(GF(1))], field=GF))
        if q > 2:
            roots_case(f"{name}/irreducible", GF, galois.irreducible_poly(q, 3))
    out["roots/names"] = np.array(sorted({k.split("/")[1] + "/" + k.split("/")[2] for k in out if k.startswith("roots/")}))

    # ---- mixed-radix DFT (fft_jit general branch): prime fields with odd / composite sizes, GF(2^m) with n | 2^m - 1
    meta["fft"] = {}
    fft_cases = [("gf31", galois.GF(31), [3, 5, 6, 10, 15, 30]),
                 ("gf998244353", galois.GF(998244353), [7, 17, 119, 14, 7 * 64, 17 * 8 * 7, 1]),
                 ("gf2147483647", galois.GF(2**31 - 1), [2, 9, 18, 63, 693, 331, 2 * 3 * 7 * 11 * 31]),
                 ("gf256", galois.GF(2**8), [3, 5, 15, 17, 51, 85, 255]),
                 ("gf4096", galois.GF(2**12), [9, 63, 65, 455, 4095]),
                 ("gf65536", galois.GF(2**16), [257, 771])]
    for name, GF, sizes in fft_cases:
        meta["fft"][name] = {"order": int(GF.order), "irreducible_poly": int(GF.irreducible_poly),
                             "primitive_element": int(GF.primitive_element), "sizes": sizes,
                             "omegas": [int(GF.primitive_root_of_unity(n)) for n in sizes]}
        for n in sizes:
            x = GF.Random(n, seed=int(rng.integers(1 << 30)))
            out[f"fft/{name}/{n}/x"] = np.asarray(x, dtype=np.int64)
            out[f"fft/{name}/{n}/X"] = np.asarray(np.fft.fft(x), dtype=np.int64)
            out[f"fft/{name}/{n}/ix"] = np.asarray(np.fft.ifft(x), dtype=np.int64)

    # ---- dense linear algebra at larger shapes, incl. singular / rectangular / ncols-limited cases
    meta["linalg"] = {}
    la_fields = {"gf256": galois.GF(2**8), "gf998244353": galois.GF(998244353), "gf2147483647": galois.GF(2**31 - 1),
                 "gf2": galois.GF(2), "gf2_20": galois.GF(2**20)}
    for name, GF in la_fields.items():
        meta["linalg"][name] = {"order": int(GF.order), "irreducible_poly": int(GF.irreducible_poly),
                                "primitive_element": int(GF.primitive_element)}
        seed = lambda: int(rng.integers(1 << 30))  # noqa: E731
        mats = {}
        mats["sq40"] = GF.Random((40, 40), seed=seed())
        mats["wide17x45"] = GF.Random((17, 45), seed=seed())
        mats["tall45x17"] = GF.Random((45, 17), seed=seed())
        low = GF.Random((30, 6), seed=seed()) @ GF.Random((6, 30), seed=seed())  # rank <= 6
        mats["rank6_30"] = low
        z = GF.Random((12, 12), seed=seed())
        z[:, 0] = 0
        z[3, :] = 0
        z[0, 1] = 0
        mats["zero_col_row"] = z
        for key, A in mats.items():
            base = f"linalg/{name}/{key}"
            out[f"{base}/A"] = np.asarray(A, dtype=np.int64)
            out[f"{base}/rre"] = np.asarray(A.row_reduce(), dtype=np.int64)
            nc = max(1, A.shape[1] // 2)
            out[f"{base}/rre_ncols"] = np.asarray(A.row_reduce(ncols=nc), dtype=np.int64)
            out[f"{base}/rre_right"] = np.asarray(A.row_reduce(eye="right"), dtype=np.int64)
            out[f"{base}/rank"] = np.array(int(np.linalg.matrix_rank(A)))
            P, L, U = A.plu_decompose()
            out[f"{base}/P"], out[f"{base}/L"], out[f"{base}/U"] = (np.asarray(v, dtype=np.int64) for v in (P, L, U))
            out[f"{base}/row_space"] = np.asarray(A.row_space(), dtype=np.int64)
            out[f"{base}/column_space"] = np.asarray(A.column_space(), dtype=np.int64)
            out[f"{base}/left_null_space"] = np.asarray(A.left_null_space(), dtype=np.int64)
            out[f"{base}/null_space"] = np.asarray(A.null_space(), dtype=np.int64)
            try:
                L2, U2 = A.lu_decompose()
                out[f"{base}/lu_L"], out[f"{base}/lu_U"] = np.asarray(L2, dtype=np.int64), np.asarray(U2, dtype=np.int64)
            except (ValueError, IndexError):
                out[f"{base}/lu_raises"] = np.array(1)
            if A.shape[0] == A.shape[1]:
                out[f"{base}/det"] = np.array(int(np.linalg.det(A)))
                try:
                    out[f"{base}/inv"] = np.asarray(np.linalg.inv(A), dtype=np.int64)
                    b1, b2 = GF.Random(A.shape[0], seed=seed()), GF.Random((A.shape[0], 3), seed=seed())
                    out[f"{base}/b1"], out[f"{base}/b2"] = np.asarray(b1, dtype=np.int64), np.asarray(b2, dtype=np.int64)
                    out[f"{base}/x1"] = np.asarray(np.linalg.solve(A, b1), dtype=np.int64)
                    out[f"{base}/x2"] = np.asarray(np.linalg.solve(A, b2), dtype=np.int64)
                except np.linalg.LinAlgError as e:
                    out[f"{base}/inv_raises"] = np.array(str(e))
        meta["linalg"][name]["cases"] = list(mats)
    np.savez_compressed(OUT / "live_next.npz", **out)
    json.dump(meta, open(OUT / "live_next_meta.json", "w"), indent=1)
    print("live_next.npz", len(out), "arrays")


def main():
    which = sys.argv[1:] or ["fields", "rs_ref", "elementwise", "gf128", "ntt", "linalg", "rs", "bch_ref", "bch",
                             "linalg_ref", "roots_ref", "next"]
    if "fields" in which:
        convert_field_vectors()
    if "rs_ref" in which:
        convert_rs_vectors()
    if "elementwise" in which:
        live_elementwise()
    if "gf128" in which:
        live_gf2_128()
    if "ntt" in which:
        live_ntt()
    if "linalg" in which:
        live_linalg()
    if "rs" in which:
        live_rs()
    if "bch_ref" in which:
        convert_bch_vectors()
    if "bch" in which:
        live_bch()
    if "linalg_ref" in which:
        convert_linalg_vectors()
    if "roots_ref" in which:
        convert_roots_vectors()
    if "next" in which:
        live_next_rows()


if __name__ == "__main__":
    main()
