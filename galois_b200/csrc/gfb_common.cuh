// gfb_common.cuh — shared device-side arithmetic for the galois_b200 kernels (sm_100a).
//
// Field "functors" hold the per-field constants in registers / kernel params and
// expose add/sub/neg/mul/inv1 on a compute element type.  The element-wise, NTT,
// convolution, matmul and Reed-Solomon kernels are templates over these functors.
//
// Reference semantics each functor reproduces (paths under /root/reference/src/galois):
//   PrimeBarrett / PrimeMersenne31 : _domains/_calculate.py:133-147 (add), :184-199 (neg),
//       :235-251 (sub), :327-340 (mul = (a*b) % p), :386-417 (inverse; any algorithm gives
//       the unique inverse, we use Fermat + Montgomery batch inversion instead of EGCD).
//   BinaryCalc   : _fields/_ufunc.py:59-61 (add/sub = xor, neg = identity),
//       _domains/_calculate.py:288-324 (shift-and-xor carry-less multiply with reduction),
//       :447-489 (inverse = a^(2^m-2)).
//   BinaryLut    : _domains/_lookup.py:143-270 (EXP/LOG arithmetic), tables built as in :319-371.
//   Binary128    : same as BinaryCalc on two 64-bit limbs (the reference runs these fields in
//       "python-calculate" mode on Python ints, _domains/_ufunc.py:146-159).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "galois_b200.h"

namespace gfb {

// ----------------------------------------------------------------------------------------
// Host-side helpers
// ----------------------------------------------------------------------------------------
void set_last_error(const char* what, cudaError_t err);
#define GFB_CUDA_TRY(expr)                              \
    do {                                                \
        cudaError_t _e = (expr);                        \
        if (_e != cudaSuccess) {                        \
            ::gfb::set_last_error(#expr, _e);           \
            return GFB_ERR_CUDA;                        \
        }                                               \
    } while (0)

int sm_count();  // multiprocessors of the current device (cached)
void conv_cache_drop(gfb_field* f);  // gfb_linalg.cu: transform plans cached for gfb_convolve over this field

// ----------------------------------------------------------------------------------------
// 128-bit element for GF(2^m), 64 < m <= 128
// ----------------------------------------------------------------------------------------
struct __align__(16) u128 {
    unsigned long long lo, hi;
};
__host__ __device__ __forceinline__ bool operator==(const u128& a, const u128& b) {
    return a.lo == b.lo && a.hi == b.hi;
}

// ----------------------------------------------------------------------------------------
// Prime fields, p < 2^32.  Compute type uint32_t.
// ----------------------------------------------------------------------------------------
struct PrimeBarrett {
    using elem = uint32_t;
    static constexpr bool kNeedsSmem = false;
    uint32_t p;
    unsigned long long mu;  // floor(2^64 / p)
    uint32_t pinv;          // p^-1 mod 2^32 (odd p), for Montgomery reduction
    uint32_t r2;            // 2^64 mod p

    __device__ __forceinline__ void init_smem(void*) const {}
    __device__ __forceinline__ elem zero() const { return 0u; }
    __device__ __forceinline__ elem one() const { return p == 1u ? 0u : 1u; }
    __device__ __forceinline__ bool is_zero(elem a) const { return a == 0u; }
    __device__ __forceinline__ elem add(elem a, elem b) const {
        uint32_t s = a + b;
        // a + b may wrap past 2^32 when p > 2^31: the wrapped sum is then < a.
        if (s < a || s >= p) s -= p;
        return s;
    }
    __device__ __forceinline__ elem sub(elem a, elem b) const {
        uint32_t d = a - b;
        if (a < b) d += p;
        return d;
    }
    __device__ __forceinline__ elem neg(elem a) const { return a ? p - a : 0u; }
    __device__ __forceinline__ elem reduce64(unsigned long long x) const {
        // Barrett: q in {floor(x/p)-1, floor(x/p)} for any x < 2^64.
        unsigned long long q = __umul64hi(x, mu);
        unsigned long long r = x - q * (unsigned long long)p;
        if (r >= p) r -= p;
        return (uint32_t)r;
    }
    // a * b * 2^-32 mod p for a, b < p (p odd): Montgomery reduction, 3 multiplies
    __device__ __forceinline__ elem montmul(elem a, elem b) const {
        const unsigned long long t = (unsigned long long)a * (unsigned long long)b;
        const uint32_t m = (uint32_t)t * pinv;
        const uint32_t u = __umulhi(m, p);
        const uint32_t hi = (uint32_t)(t >> 32);
        uint32_t r = hi - u;  // t - m*p is divisible by 2^32: the low words cancel exactly
        if (hi < u) r += p;
        return r;
    }
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        // two Montgomery steps: (a b 2^-32) (2^64) 2^-32 = a b.  ~12 integer instructions instead of a 64-bit Barrett.
        if (p == 2u) return a & b;
        return montmul(montmul(a, b), r2);
    }
    // order - 1 as an unsigned 64-bit integer (exponent modulus)
    __device__ __forceinline__ unsigned long long group_order() const { return (unsigned long long)p - 1ull; }
    // square-and-multiply inside the Montgomery domain: one montmul per squaring / multiply instead of two
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        if (p == 2u) return e == 0ull ? 1u : a;
        elem x = montmul(a, r2);   // a R
        elem r = montmul(1u, r2);  // R  (one)
        while (e) {
            if (e & 1ull) r = montmul(r, x);
            e >>= 1;
            if (e) x = montmul(x, x);
        }
        return montmul(r, 1u);
    }
    // inverse of a non-zero element (Fermat: a^(p-2)).  Branch-free, fixed 32 rounds in the Montgomery domain: the
    // multiply chain (r) and the squaring chain (x) are independent, so a lone thread — the pivot inversion of the
    // Gaussian elimination — sees ~32 x one Montgomery latency instead of a data-dependent loop with two branches
    // per exponent bit.  Squarings past the top bit are harmless.
    __device__ __forceinline__ elem inv1(elem a) const {
        if (p == 2u) return a;
        elem x = montmul(a, r2);   // a R
        elem r = montmul(1u, r2);  // R
        uint32_t e = p - 2u;
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const elem t = montmul(r, x);
            r = (e & 1u) ? t : r;
            x = montmul(x, x);
            e >>= 1;
        }
        return montmul(r, 1u);
    }
    // a multiplier used many times: keep it in Montgomery form so that each product is ONE montmul
    __device__ __forceinline__ elem premul(elem b) const { return p == 2u ? b : montmul(b, r2); }
    __device__ __forceinline__ elem mul_prepared(elem a, elem b_prepared) const {
        return p == 2u ? (a & b_prepared) : montmul(a, b_prepared);
    }
};

struct PrimeMersenne31 {
    using elem = uint32_t;
    static constexpr bool kNeedsSmem = false;
    static constexpr uint32_t p = 0x7fffffffu;

    __device__ __forceinline__ void init_smem(void*) const {}
    __device__ __forceinline__ elem zero() const { return 0u; }
    __device__ __forceinline__ elem one() const { return 1u; }
    __device__ __forceinline__ bool is_zero(elem a) const { return a == 0u; }
    __device__ __forceinline__ elem add(elem a, elem b) const {
        uint32_t s = a + b;  // < 2^32
        if (s >= p) s -= p;
        return s;
    }
    __device__ __forceinline__ elem sub(elem a, elem b) const {
        uint32_t d = a - b;
        if (a < b) d += p;
        return d;
    }
    __device__ __forceinline__ elem neg(elem a) const { return a ? p - a : 0u; }
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        unsigned long long x = (unsigned long long)a * (unsigned long long)b;  // < 2^62
        uint32_t lo = (uint32_t)x & p;
        uint32_t hi = (uint32_t)(x >> 31);  // < 2^31
        uint32_t s = lo + hi;               // < 2^32
        if (s >= p) s -= p;
        return s;
    }
    __device__ __forceinline__ unsigned long long group_order() const { return (unsigned long long)p - 1ull; }
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        elem r = 1u;
        while (e) {
            if (e & 1ull) r = mul(r, a);
            e >>= 1;
            if (e) a = mul(a, a);
        }
        return r;
    }
    __device__ __forceinline__ elem inv1(elem a) const {
        // a^(2^31 - 3): exponent bits are 29 ones, a zero, a one  → fixed addition chain.
        return pow_u64(a, (unsigned long long)p - 2ull);
    }
    __device__ __forceinline__ elem premul(elem b) const { return b; }
    __device__ __forceinline__ elem mul_prepared(elem a, elem b_prepared) const { return mul(a, b_prepared); }
};

// ----------------------------------------------------------------------------------------
// Carry-less multiplication without a CLMUL instruction: operands are split into four bit-slices
// (bits = i mod 4) so that the integer multiplier's carries land in the three-zero gaps; each slice
// has at most 8 set bits, so a column sum is at most 8 and never reaches the next slice position.
// 16 IMAD.WIDE + ~30 logic operations per 32x32 -> 64 bit product.
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mul_wide_u32(uint32_t a, uint32_t b) {
    unsigned long long r;  // one IMAD.WIDE.U32 (the C++ form leaves dead 64x64 cross-term adds behind)
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}

__device__ __forceinline__ unsigned long long clmul32(uint32_t x, uint32_t y) {
    const uint32_t x0 = x & 0x11111111u, x1 = x & 0x22222222u, x2 = x & 0x44444444u, x3 = x & 0x88888888u;
    const uint32_t y0 = y & 0x11111111u, y1 = y & 0x22222222u, y2 = y & 0x44444444u, y3 = y & 0x88888888u;
#define GFB_M(a, b) mul_wide_u32(a, b)
    unsigned long long z0 = GFB_M(x0, y0) ^ GFB_M(x1, y3) ^ GFB_M(x2, y2) ^ GFB_M(x3, y1);
    unsigned long long z1 = GFB_M(x0, y1) ^ GFB_M(x1, y0) ^ GFB_M(x2, y3) ^ GFB_M(x3, y2);
    unsigned long long z2 = GFB_M(x0, y2) ^ GFB_M(x1, y1) ^ GFB_M(x2, y0) ^ GFB_M(x3, y3);
    unsigned long long z3 = GFB_M(x0, y3) ^ GFB_M(x1, y2) ^ GFB_M(x2, y1) ^ GFB_M(x3, y0);
#undef GFB_M
    return (z0 & 0x1111111111111111ull) | (z1 & 0x2222222222222222ull) | (z2 & 0x4444444444444444ull) |
           (z3 & 0x8888888888888888ull);
}

// 64 x 64 -> 128 (Karatsuba over 32-bit halves: 3 clmul32)
__device__ __forceinline__ void clmul64(unsigned long long a, unsigned long long b, unsigned long long& lo,
                                        unsigned long long& hi) {
    const uint32_t a0 = (uint32_t)a, a1 = (uint32_t)(a >> 32), b0 = (uint32_t)b, b1 = (uint32_t)(b >> 32);
    const unsigned long long p0 = clmul32(a0, b0), p2 = clmul32(a1, b1);
    const unsigned long long p1 = clmul32(a0 ^ a1, b0 ^ b1) ^ p0 ^ p2;
    lo = p0 ^ (p1 << 32);
    hi = p2 ^ (p1 >> 32);
}

// spread the 32 bits of x to the even bit positions of a 64-bit word (squaring in GF(2)[x])
__device__ __forceinline__ unsigned long long spread32(uint32_t x) {
    unsigned long long v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// ----------------------------------------------------------------------------------------
// Binary extension fields GF(2^m), 1 <= m <= 63, explicit calculation.  Compute type uint64_t.
// The product is the same element the reference's shift-and-xor loop yields (_calculate.py:309-324):
// full carry-less product, then the part above x^m is folded back with r(x) = poly - x^m.
// ----------------------------------------------------------------------------------------
struct BinaryCalc {
    using elem = unsigned long long;
    static constexpr bool kNeedsSmem = false;
    int m;
    unsigned long long poly_low;  // r(x): irreducible polynomial WITHOUT the x^m term
    unsigned long long mask;      // 2^m - 1

    __device__ __forceinline__ void init_smem(void*) const {}
    __device__ __forceinline__ elem zero() const { return 0ull; }
    __device__ __forceinline__ elem one() const { return 1ull; }
    __device__ __forceinline__ bool is_zero(elem a) const { return a == 0ull; }
    __device__ __forceinline__ elem add(elem a, elem b) const { return a ^ b; }
    __device__ __forceinline__ elem sub(elem a, elem b) const { return a ^ b; }
    __device__ __forceinline__ elem neg(elem a) const { return a; }
    // (c1:c0) mod x^m + r(x); the degree drops by m - deg r >= 1 per fold
    __device__ __forceinline__ elem reduce(unsigned long long c0, unsigned long long c1) const {
        if (m <= 32) {
            // c1 == 0 and every intermediate fits 64 bits
#pragma unroll 1
            for (int guard = 0; guard < 64; ++guard) {
                const unsigned long long h = c0 >> m;
                if (h == 0ull) break;
                c0 = (c0 & mask) ^ clmul32((uint32_t)h, (uint32_t)poly_low);
            }
            return c0;
        }
#pragma unroll 1
        for (int guard = 0; guard < 128; ++guard) {
            const unsigned long long h = (c0 >> m) | (c1 << (64 - m));  // c has < 2m bits: h fits 64 bits
            if (h == 0ull) break;
            unsigned long long t0, t1;
            clmul64(h, poly_low, t0, t1);
            c0 = (c0 & mask) ^ t0;
            c1 = t1;
        }
        return c0 & mask;
    }
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        if (m <= 32) return reduce(clmul32((uint32_t)a, (uint32_t)b), 0ull);
        unsigned long long c0, c1;
        clmul64(a, b, c0, c1);
        return reduce(c0, c1);
    }
    __device__ __forceinline__ elem sqr(elem a) const {
        if (m <= 32) return reduce(spread32((uint32_t)a), 0ull);
        return reduce(spread32((uint32_t)a), spread32((uint32_t)(a >> 32)));
    }
    __device__ __forceinline__ unsigned long long group_order() const { return mask; }  // 2^m - 1
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        elem r = 1ull;
        while (e) {
            if (e & 1ull) r = mul(r, a);
            e >>= 1;
            if (e) a = sqr(a);
        }
        return r;
    }
    // a^-1 = (a^(2^(m-1) - 1))^2, Itoh-Tsujii chain (see Binary128::inv1)
    __device__ __forceinline__ elem inv1(elem a) const {
        if (m == 1) return a;
        const int e = m - 1;
        const int top = 31 - __clz(e);
        elem beta = a;
        int k = 1;
#pragma unroll 1
        for (int bit = top - 1; bit >= 0; --bit) {
            elem t = beta;
#pragma unroll 1
            for (int i = 0; i < k; ++i) t = sqr(t);
            beta = mul(t, beta);
            k <<= 1;
            if ((e >> bit) & 1) {
                beta = mul(sqr(beta), a);
                k += 1;
            }
        }
        return sqr(beta);
    }
    __device__ __forceinline__ elem premul(elem b) const { return b; }
    __device__ __forceinline__ elem mul_prepared(elem a, elem b_prepared) const { return mul(a, b_prepared); }
};

// ----------------------------------------------------------------------------------------
// GF(2^m), m <= 12, EXP/LOG tables staged in shared memory.  Compute type uint32_t.
// Shared layout: uint16 EXP[2*order] followed by uint16 LOG[order].
// ----------------------------------------------------------------------------------------
struct BinaryLut {
    using elem = uint32_t;
    static constexpr bool kNeedsSmem = true;
    int m;
    uint32_t order;               // 2^m
    const uint16_t* g_exp;        // device global, 2*order entries
    const uint16_t* g_log;        // device global, order entries
    // set by init_smem (per-thread copies of the shared pointers)
    const uint16_t* s_exp;
    const uint16_t* s_log;

    __host__ __device__ size_t smem_bytes() const { return (size_t)order * 3u * sizeof(uint16_t); }
    __device__ __forceinline__ void init_smem(void* smem) {
        uint16_t* e = reinterpret_cast<uint16_t*>(smem);
        uint16_t* l = e + 2u * order;
        for (uint32_t i = threadIdx.x; i < 2u * order; i += blockDim.x) e[i] = g_exp[i];
        for (uint32_t i = threadIdx.x; i < order; i += blockDim.x) l[i] = g_log[i];
        s_exp = e;
        s_log = l;
        __syncthreads();
    }
    __device__ __forceinline__ elem zero() const { return 0u; }
    __device__ __forceinline__ elem one() const { return 1u; }
    __device__ __forceinline__ bool is_zero(elem a) const { return a == 0u; }
    __device__ __forceinline__ elem add(elem a, elem b) const { return a ^ b; }
    __device__ __forceinline__ elem sub(elem a, elem b) const { return a ^ b; }
    __device__ __forceinline__ elem neg(elem a) const { return a; }
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        if (a == 0u || b == 0u) return 0u;
        return s_exp[(uint32_t)s_log[a] + (uint32_t)s_log[b]];
    }
    __device__ __forceinline__ unsigned long long group_order() const { return (unsigned long long)order - 1ull; }
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        // e already reduced mod (order-1) by the caller; a != 0
        if (e == 0ull) return 1u;
        unsigned long long idx = ((unsigned long long)s_log[a] * e) % ((unsigned long long)order - 1ull);
        return s_exp[idx];
    }
    __device__ __forceinline__ elem inv1(elem a) const { return s_exp[(order - 1u) - (uint32_t)s_log[a]]; }
    __device__ __forceinline__ elem premul(elem b) const { return b; }
    __device__ __forceinline__ elem mul_prepared(elem a, elem b_prepared) const { return mul(a, b_prepared); }
};

// ----------------------------------------------------------------------------------------
// GF(2^m), 64 <= m <= 128 on two 64-bit limbs.  Multiply = 128x128 Karatsuba (9 clmul32) followed by
// folding the high part with the low-degree tail r(x) of the irreducible polynomial x^m + r(x).
// Inverse = Itoh-Tsujii addition chain on 2^(m-1) - 1 (about log2(m) + popcount(m-1) multiplies and
// m - 1 squarings) instead of the reference's plain square-and-multiply of a^(2^m - 2)
// (_calculate.py:470-489) — the inverse is unique, so the result is the same element.
// ----------------------------------------------------------------------------------------
struct Binary128 {
    using elem = u128;
    static constexpr bool kNeedsSmem = false;
    int m;
    u128 poly_low;  // r(x): irreducible polynomial without the x^m term
    u128 mask;      // 2^m - 1
    int n_shift;    // > 0: r(x) = sum x^shift[i] with every shift < 64 (sparse tail, e.g. GCM's x^7+x^2+x+1)
    unsigned char shift[8];

    __device__ __forceinline__ void init_smem(void*) const {}
    __device__ __forceinline__ elem zero() const { return u128{0ull, 0ull}; }
    __device__ __forceinline__ elem one() const { return u128{1ull, 0ull}; }
    __device__ __forceinline__ bool is_zero(elem a) const { return (a.lo | a.hi) == 0ull; }
    __device__ __forceinline__ elem add(elem a, elem b) const { return u128{a.lo ^ b.lo, a.hi ^ b.hi}; }
    __device__ __forceinline__ elem sub(elem a, elem b) const { return add(a, b); }
    __device__ __forceinline__ elem neg(elem a) const { return a; }

    // 128 x 128 -> 256 bits, c[0] least significant
    static __device__ __forceinline__ void mul256(const u128& a, const u128& b, unsigned long long (&c)[4]) {
        unsigned long long p0, p1, q0, q1, m0, m1;
        clmul64(a.lo, b.lo, p0, p1);
        clmul64(a.hi, b.hi, q0, q1);
        clmul64(a.lo ^ a.hi, b.lo ^ b.hi, m0, m1);
        m0 ^= p0 ^ q0;
        m1 ^= p1 ^ q1;
        c[0] = p0;
        c[1] = p1 ^ m0;
        c[2] = q0 ^ m1;
        c[3] = q1;
    }

    // c (256 bits) mod x^m + r(x)
    __device__ __forceinline__ elem reduce(unsigned long long (&c)[4]) const {
        const int sh = m - 64;  // 0..64
#pragma unroll 1
        for (int guard = 0; guard < 256; ++guard) {
            // H = c >> m  (at most 128 bits), L = c mod x^m
            unsigned long long h0, h1;
            if (sh == 0) {
                h0 = c[1];
                h1 = c[2];
            } else if (sh == 64) {
                h0 = c[2];
                h1 = c[3];
            } else {
                h0 = (c[1] >> sh) | (c[2] << (64 - sh));
                h1 = (c[2] >> sh) | (c[3] << (64 - sh));
            }
            if ((h0 | h1) == 0ull) break;
            const u128 L{c[0] & mask.lo, c[1] & mask.hi};
            if (n_shift > 0) {
                unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull;
                for (int i = 0; i < n_shift; ++i) {
                    const int s = shift[i];
                    if (s == 0) {
                        t0 ^= h0;
                        t1 ^= h1;
                    } else {
                        t0 ^= h0 << s;
                        t1 ^= (h1 << s) | (h0 >> (64 - s));
                        t2 ^= h1 >> (64 - s);
                    }
                }
                c[0] = L.lo ^ t0;
                c[1] = L.hi ^ t1;
                c[2] = t2;
                c[3] = 0ull;
            } else {
                unsigned long long t[4];
                mul256(u128{h0, h1}, poly_low, t);
                c[0] = L.lo ^ t[0];
                c[1] = L.hi ^ t[1];
                c[2] = t[2];
                c[3] = t[3];
            }
        }
        return u128{c[0] & mask.lo, c[1] & mask.hi};
    }

    __device__ __forceinline__ elem mul(elem a, elem b) const {
        unsigned long long c[4];
        mul256(a, b, c);
        return reduce(c);
    }
    __device__ __forceinline__ elem sqr(elem a) const {
        unsigned long long c[4];
        c[0] = spread32((uint32_t)a.lo);
        c[1] = spread32((uint32_t)(a.lo >> 32));
        c[2] = spread32((uint32_t)a.hi);
        c[3] = spread32((uint32_t)(a.hi >> 32));
        return reduce(c);
    }
    __device__ __forceinline__ unsigned long long group_order() const { return ~0ull; }  // not used (> 2^64)
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        elem r = one();
        while (e) {
            if (e & 1ull) r = mul(r, a);
            e >>= 1;
            if (e) a = sqr(a);
        }
        return r;
    }
    // a^-1 = (a^(2^(m-1) - 1))^2 with beta_k = a^(2^k - 1):  beta_2k = beta_k^(2^k) * beta_k,  beta_(k+1) = beta_k^2 * a
    __device__ __forceinline__ elem inv1(elem a) const {
        const int e = m - 1;
        int top = 31 - __clz(e);
        elem beta = a;
        int k = 1;
#pragma unroll 1
        for (int bit = top - 1; bit >= 0; --bit) {
            elem t = beta;
#pragma unroll 1
            for (int i = 0; i < k; ++i) t = sqr(t);
            beta = mul(t, beta);
            k <<= 1;
            if ((e >> bit) & 1) {
                beta = mul(sqr(beta), a);
                k += 1;
            }
        }
        return sqr(beta);
    }
    __device__ __forceinline__ elem premul(elem b) const { return b; }
    __device__ __forceinline__ elem mul_prepared(elem a, elem b_prepared) const { return mul(a, b_prepared); }
};

// ----------------------------------------------------------------------------------------
// Storage <-> compute conversions
// ----------------------------------------------------------------------------------------
template <typename E, typename T>
__device__ __forceinline__ E to_elem(T v) {
    return (E)v;
}
template <>
__device__ __forceinline__ u128 to_elem<u128, u128>(u128 v) {
    return v;
}
template <typename T, typename E>
__device__ __forceinline__ T to_store(E v) {
    return (T)v;
}
template <>
__device__ __forceinline__ u128 to_store<u128, u128>(u128 v) {
    return v;
}

// 128-bit streaming global access (evict-first: these arrays are touched once)
template <typename T>
struct alignas(16) Pack {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};
template <typename T>
__device__ __forceinline__ Pack<T> ld_stream(const T* p) {
    uint4 r = __ldcs(reinterpret_cast<const uint4*>(p));
    return *reinterpret_cast<Pack<T>*>(&r);
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, const Pack<T>& v) {
    __stcs(reinterpret_cast<uint4*>(p), *reinterpret_cast<const uint4*>(&v));
}

// ----------------------------------------------------------------------------------------
// The opaque field handle (host side)
// ----------------------------------------------------------------------------------------
enum FieldImpl { IMPL_PRIME_BARRETT = 0, IMPL_PRIME_M31 = 1, IMPL_BIN_LUT = 2, IMPL_BIN_CALC = 3, IMPL_BIN_128 = 4 };

}  // namespace gfb

struct gfb_field {
    int kind;
    int impl;    // gfb::FieldImpl
    int device;  // CUDA device the tables live on
    int degree;
    uint64_t order_lo;       // p, or 2^m when m < 64 (0 otherwise)
    uint64_t poly_lo, poly_hi;
    uint64_t primitive_element;
    // Barrett / Montgomery constants
    uint32_t p;
    unsigned long long mu;
    uint32_t pinv, r2;
    // lookup tables (m <= 16 built on host; staged in smem when m <= 12)
    uint16_t* d_exp;
    uint16_t* d_log;
    uint16_t* h_exp;
    uint16_t* h_log;
    // 256 x 256 product + quotient tables of a byte-sized GF(2^m) (gfb_ewise_tab8.cu), built on first use
    uint8_t* d_tab8;
    // sticky error flag in mapped pinned host memory
    int* h_err;
    int* d_err;

    gfb::PrimeBarrett prime_barrett() const { return gfb::PrimeBarrett{p, mu, pinv, r2}; }
    gfb::PrimeMersenne31 prime_m31() const { return gfb::PrimeMersenne31{}; }
    gfb::BinaryCalc bin_calc() const {
        unsigned long long mask = degree == 64 ? ~0ull : ((1ull << degree) - 1ull);
        return gfb::BinaryCalc{degree, poly_lo & mask, mask};
    }
    gfb::BinaryLut bin_lut() const {
        return gfb::BinaryLut{degree, 1u << degree, d_exp, d_log, nullptr, nullptr};
    }
    gfb::Binary128 bin_128() const {
        gfb::u128 mask, pl;
        if (degree == 128) {
            mask = gfb::u128{~0ull, ~0ull};
        } else {
            mask = gfb::u128{~0ull, (1ull << (degree - 64)) - 1ull};
        }
        pl = gfb::u128{poly_lo & mask.lo, poly_hi & mask.hi};
        gfb::Binary128 b{degree, pl, mask, 0, {0, 0, 0, 0, 0, 0, 0, 0}};
        // sparse low-degree tail: fold the high half with a few shifts instead of another 128x128 product
        if (pl.hi == 0ull && __builtin_popcountll(pl.lo) <= 8) {
            for (int i = 0; i < 64; ++i)
                if ((pl.lo >> i) & 1ull) b.shift[b.n_shift++] = (unsigned char)i;
        }
        return b;
    }
};
