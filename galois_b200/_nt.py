"""
Host-side number theory needed to *describe* a field to the CUDA backend (cold path, run once per
field): primality, factoring, smallest primitive root, roots of unity, GF(2)[x] irreducibility.

These mirror what the reference computes on the host before any kernel runs
(/root/reference/src/galois/_prime.py, _primitive_root.py:124 "min" primitive root,
_fields/_array.py:1127-1187 primitive_root_of_unity, _fields/_factory.py:364-535) so that the
constants handed to the kernels (g, omega, irreducible polynomial, primitive element) are the
reference's constants.  SURVEY.md §2 rows 12-13 mark this as one-off host setup.
"""

from __future__ import annotations

import functools
import math
import random

# Conway polynomials C_{2,m} as integers (bit i = coefficient of x^i); the reference's default
# irreducible polynomial for GF(2^m) (_fields/_factory.py:447-450, conway_polys.db).  x (= 2) is a
# primitive element for each of them.  Degrees absent here (e.g. 128) raise LookupError in the
# reference as well.
CONWAY_2 = {
    1: 0x3, 2: 0x7, 3: 0xB, 4: 0x13, 5: 0x25, 6: 0x5B, 7: 0x83, 8: 0x11D, 9: 0x211, 10: 0x46F, 11: 0x805,
    12: 0x10EB, 13: 0x201B, 14: 0x40A9, 15: 0x8035, 16: 0x1002D, 17: 0x20009, 18: 0x41403, 19: 0x80027,
    20: 0x1006F3, 21: 0x200065, 22: 0x401F61, 23: 0x800021, 24: 0x101E6A9, 25: 0x2000145, 26: 0x40045D3,
    27: 0x80016AD, 28: 0x100020E5, 29: 0x20000005, 30: 0x400328AF, 31: 0x80000009, 32: 0x100008299,
    33: 0x200003D49, 34: 0x4000199F7, 35: 0x800000CA5, 36: 0x1000DA6163, 37: 0x200000003F, 38: 0x4000004727,
    39: 0x8000009EE5, 40: 0x10000A5B12B, 41: 0x20000000009, 42: 0x40047141A67, 43: 0x80000000059,
    44: 0x1000010B001B, 45: 0x20000012D841, 46: 0x400000B24001, 47: 0x800000000021, 48: 0x1000002821D89,
    49: 0x200000000055F, 50: 0x40000380B7755, 51: 0x8000000019241, 52: 0x1000001EA2C493, 53: 0x20000000000047,
    54: 0x400005EA27A097, 55: 0x80000000000E91, 56: 0x100000244486B1D, 57: 0x200000000292D7F,
    58: 0x4000000A7451DEB, 59: 0x80000000000007B, 60: 0x10003697464A113D, 61: 0x2000000000000027,
    62: 0x400000017F3F7043, 63: 0x8000000001C38B1F, 64: 0x10000000247F43CB7, 65: 0x2000000000000F6A3,
    66: 0x400006547E9E928F5, 67: 0x80000000000000027, 68: 0x10000000087450E333, 69: 0x200000000001064731,
    70: 0x40000018BF6538FF8D, 71: 0x80000000000000002B, 72: 0x10000019049EAE49111, 73: 0x200000000000000001D,
    74: 0x4000000003FFD233909, 75: 0x8000000000122C05233, 76: 0x10000000007EAB98C027, 77: 0x2000000000000000C815,
    78: 0x400000005CAF1B6EEBD3, 79: 0x8000000000000000001D, 80: 0x1000000014643A772AB75,
    81: 0x20000000000000DA38ED9, 82: 0x400000000003F1163BADF, 83: 0x800000000000000000095,
    84: 0x1000000211C2A197333BB5, 85: 0x200000000000000055C0C5, 86: 0x400000000009408A6157E7,
    87: 0x8000000000000054117DAB, 88: 0x10000000000B252D92CFBBB, 89: 0x20000000000000000000069,
    90: 0x4000001F7ADF5F557055C95, 91: 0x8000000000000000003B613, 92: 0x1000000000011F6341E7D235,
    95: 0x80000000000000000127734B, 96: 0x100000000B844512311BDEBDD, 97: 0x2000000000000000000000041,
    98: 0x400000000016D50809C44CD4B, 100: 0x10000000000391FA3CCB598B69, 101: 0x200000000000000000000000C3,
    102: 0x40000000102FDAB2A05B810401, 103: 0x800000000000000000000000BD, 107: 0x8000000000000000000000000AF,
    108: 0x100000000050EA44FC6FE5948F8D, 109: 0x2000000000000000000000000035,
    110: 0x4000000000134FC10CBA36126C47, 113: 0x2000000000000000000000000002D,
    114: 0x4000000002AF3D3D4853CDE35AA89, 115: 0x8000000000000000000000246FED3,
    119: 0x80000000000000000000000037E087, 120: 0x1000000006A8A9A151DF754F20370BD,
    121: 0x2000000000000000000000000003CC5, 125: 0x200000000000000000000000004BBE69,
    126: 0x4000000002E2A4D15B18BFB0135BAD63, 127: 0x80000000000000000000000000000003,
}

_SMALL_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n: int) -> bool:
    """Deterministic Miller-Rabin for n < 3.3e24, probabilistic (40 rounds) beyond."""
    if n < 2:
        return False
    for q in _SMALL_PRIMES:
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    bases = _SMALL_PRIMES if n < 3_317_044_064_679_887_385_961_981 else [random.randrange(2, n - 1) for _ in range(40)]
    for a in bases:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _pollard_rho(n: int) -> int:
    if n % 2 == 0:
        return 2
    rng = random.Random(n)
    while True:
        c = rng.randrange(1, n)
        f = lambda v: (v * v + c) % n  # noqa: E731
        x = y = rng.randrange(2, n)
        d = 1
        while d == 1:
            x = f(x)
            y = f(f(y))
            d = math.gcd(abs(x - y), n)
        if d != n:
            return d


@functools.lru_cache(maxsize=None)
def factors(n: int) -> tuple[tuple[int, ...], tuple[int, ...]]:
    """(primes ascending, multiplicities) of n >= 1."""
    out: dict[int, int] = {}

    def rec(v: int):
        if v == 1:
            return
        if is_prime(v):
            out[v] = out.get(v, 0) + 1
            return
        for q in range(2, 1000):
            if v % q == 0:
                out[q] = out.get(q, 0) + 1
                rec(v // q)
                return
        d = _pollard_rho(v)
        rec(d)
        rec(v // d)

    rec(n)
    primes = tuple(sorted(out))
    return primes, tuple(out[q] for q in primes)


def prime_factors_with_multiplicity(n: int) -> list[int]:
    """176 → [2, 2, 2, 2, 11]  (fft_jit._prime_factors, _domains/_function.py:214-229)."""
    if n == 1:
        return []
    primes, mult = factors(n)
    return [q for q, k in zip(primes, mult) for _ in range(k)]


def ilog(n: int, b: int) -> int:
    """floor(log_b(n))."""
    k = 0
    while n >= b:
        n //= b
        k += 1
    return k


def prime_power(order: int) -> tuple[int, int]:
    """order = p^m → (p, m); ValueError otherwise."""
    if order < 2:
        raise ValueError(f"Argument 'order' must be a prime power, not {order}.")
    primes, mult = factors(order)
    if len(primes) != 1:
        raise ValueError(f"Argument 'order' must be a prime power, not {order}.")
    return primes[0], mult[0]


@functools.lru_cache(maxsize=None)
def primitive_root(p: int) -> int:
    """Smallest primitive root modulo the prime p (the reference's default, _primitive_root.py:124)."""
    if p == 2:
        return 1
    primes, _ = factors(p - 1)
    for g in range(2, p):
        if all(pow(g, (p - 1) // q, p) != 1 for q in primes):
            return g
    raise RuntimeError("no primitive root found")


def is_primitive_root(g: int, p: int) -> bool:
    if p == 2:
        return g % 2 == 1
    primes, _ = factors(p - 1)
    return g % p != 0 and all(pow(g, (p - 1) // q, p) != 1 for q in primes)


# ------------------------------------------------------------------------------- GF(2)[x] helpers
def clmul(a: int, b: int) -> int:
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        b >>= 1
    return r


def polymod2(a: int, mod: int) -> int:
    dm = mod.bit_length() - 1
    while a.bit_length() - 1 >= dm and a:
        a ^= mod << (a.bit_length() - 1 - dm)
    return a


def mulmod2(a: int, b: int, mod: int) -> int:
    return polymod2(clmul(a, b), mod)


def powmod2(a: int, e: int, mod: int) -> int:
    r = 1
    a = polymod2(a, mod)
    while e:
        if e & 1:
            r = mulmod2(r, a, mod)
        a = mulmod2(a, a, mod)
        e >>= 1
    return r


def polygcd2(a: int, b: int) -> int:
    while b:
        a, b = b, polymod2(a, b)
    return a


def is_irreducible2(poly: int) -> bool:
    """Rabin's test over GF(2): x^(2^m) == x mod f and gcd(x^(2^(m/q)) - x, f) == 1 for prime q | m."""
    m = poly.bit_length() - 1
    if m < 1:
        return False
    if m == 1:
        return True
    if poly & 1 == 0:
        return False
    x = 2
    primes, _ = factors(m)
    for q in primes:
        h = x
        for _ in range(m // q):
            h = mulmod2(h, h, poly)
        if polygcd2(h ^ x, poly) != 1:
            return False
    h = x
    for _ in range(m):
        h = mulmod2(h, h, poly)
    return h == polymod2(x, poly)


def is_primitive_element2(g: int, poly: int) -> bool:
    """Is g a multiplicative generator of GF(2)[x]/(poly)?"""
    m = poly.bit_length() - 1
    order = (1 << m) - 1
    if g == 0 or g >> m:
        return False
    if order == 1:
        return g == 1
    primes, _ = factors(order)
    return all(powmod2(g, order // q, poly) != 1 for q in primes)


def smallest_primitive_element2(poly: int) -> int:
    m = poly.bit_length() - 1
    for g in range(1 if m == 1 else 2, 1 << m):
        if is_primitive_element2(g, poly):
            return g
    raise RuntimeError("no primitive element found")


def parse_poly2(poly) -> int:
    """Accepts an int, a degree-descending coefficient list, or a string like 'x^8 + x^4 + x^3 + x^2 + 1'."""
    if isinstance(poly, int):
        return poly
    if isinstance(poly, str):
        v = 0
        for term in poly.replace(" ", "").replace("-", "+").split("+"):
            if not term:
                continue
            if term in ("1",):
                v ^= 1
            elif term == "x":
                v ^= 2
            elif term.startswith("x^") or term.startswith("x**"):
                v ^= 1 << int(term.split("^")[-1].split("*")[-1])
            elif term == "0":
                pass
            else:
                raise ValueError(f"cannot parse GF(2)[x] term {term!r}")
        return v
    v = 0
    for c in poly:
        v = (v << 1) | (int(c) & 1)
    return v
