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
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        return reduce64((unsigned long long)a * (unsigned long long)b);
    }
    // order - 1 as an unsigned 64-bit integer (exponent modulus)
    __device__ __forceinline__ unsigned long long group_order() const { return (unsigned long long)p - 1ull; }
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        elem r = one();
        while (e) {
            if (e & 1ull) r = mul(r, a);
            e >>= 1;
            if (e) a = mul(a, a);
        }
        return r;
    }
    // inverse of a non-zero element (Fermat: a^(p-2))
    __device__ __forceinline__ elem inv1(elem a) const { return pow_u64(a, (unsigned long long)p - 2ull); }
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
};

// ----------------------------------------------------------------------------------------
// Binary extension fields GF(2^m), 1 <= m <= 64, explicit calculation.  Compute type uint64_t.
// ----------------------------------------------------------------------------------------
struct BinaryCalc {
    using elem = unsigned long long;
    static constexpr bool kNeedsSmem = false;
    int m;
    unsigned long long poly_low;  // irreducible polynomial WITHOUT the x^m term
    unsigned long long mask;      // 2^m - 1

    __device__ __forceinline__ void init_smem(void*) const {}
    __device__ __forceinline__ elem zero() const { return 0ull; }
    __device__ __forceinline__ elem one() const { return 1ull; }
    __device__ __forceinline__ bool is_zero(elem a) const { return a == 0ull; }
    __device__ __forceinline__ elem add(elem a, elem b) const { return a ^ b; }
    __device__ __forceinline__ elem sub(elem a, elem b) const { return a ^ b; }
    __device__ __forceinline__ elem neg(elem a) const { return a; }
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        elem c = 0ull;
        const int top = m - 1;
        while (b) {
            if (b & 1ull) c ^= a;
            b >>= 1;
            unsigned long long carry = (a >> top) & 1ull;
            a = (a << 1) & mask;
            if (carry) a ^= poly_low;
        }
        return c;
    }
    __device__ __forceinline__ unsigned long long group_order() const { return mask; }  // 2^m - 1
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        elem r = 1ull;
        while (e) {
            if (e & 1ull) r = mul(r, a);
            e >>= 1;
            if (e) a = mul(a, a);
        }
        return r;
    }
    __device__ __forceinline__ elem inv1(elem a) const {
        // a^(2^m - 2) = prod_{i=1}^{m-1} a^(2^i)
        elem t = a, r = 1ull;
        for (int i = 1; i < m; ++i) {
            t = mul(t, t);
            r = mul(r, t);
        }
        return r;
    }
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
};

// ----------------------------------------------------------------------------------------
// GF(2^m), 64 < m <= 128 on two limbs.
// ----------------------------------------------------------------------------------------
struct Binary128 {
    using elem = u128;
    static constexpr bool kNeedsSmem = false;
    int m;
    u128 poly_low;  // irreducible polynomial without the x^m term
    u128 mask;      // 2^m - 1

    __device__ __forceinline__ void init_smem(void*) const {}
    __device__ __forceinline__ elem zero() const { return u128{0ull, 0ull}; }
    __device__ __forceinline__ elem one() const { return u128{1ull, 0ull}; }
    __device__ __forceinline__ bool is_zero(elem a) const { return (a.lo | a.hi) == 0ull; }
    __device__ __forceinline__ elem add(elem a, elem b) const { return u128{a.lo ^ b.lo, a.hi ^ b.hi}; }
    __device__ __forceinline__ elem sub(elem a, elem b) const { return add(a, b); }
    __device__ __forceinline__ elem neg(elem a) const { return a; }
    __device__ __forceinline__ elem xtime(elem a) const {
        // a * x mod poly
        const int top = m - 1;
        unsigned long long carry = top >= 64 ? (a.hi >> (top - 64)) & 1ull : (a.lo >> top) & 1ull;
        elem r;
        r.hi = ((a.hi << 1) | (a.lo >> 63)) & mask.hi;
        r.lo = (a.lo << 1) & mask.lo;
        // branch-free conditional xor
        unsigned long long sel = 0ull - carry;
        r.lo ^= poly_low.lo & sel;
        r.hi ^= poly_low.hi & sel;
        return r;
    }
    __device__ __forceinline__ elem mul(elem a, elem b) const {
        elem c = zero();
#pragma unroll 1
        for (int limb = 0; limb < 2; ++limb) {
            unsigned long long w = limb ? b.hi : b.lo;
            if (limb == 1 && w == 0ull) break;
#pragma unroll 4
            for (int i = 0; i < 64; ++i) {
                unsigned long long sel = 0ull - (w & 1ull);
                c.lo ^= a.lo & sel;
                c.hi ^= a.hi & sel;
                w >>= 1;
                a = xtime(a);
            }
        }
        return c;
    }
    __device__ __forceinline__ unsigned long long group_order() const { return ~0ull; }  // not used (> 2^64)
    __device__ __forceinline__ elem pow_u64(elem a, unsigned long long e) const {
        elem r = one();
        while (e) {
            if (e & 1ull) r = mul(r, a);
            e >>= 1;
            if (e) a = mul(a, a);
        }
        return r;
    }
    __device__ __forceinline__ elem inv1(elem a) const {
        elem t = a, r = one();
        for (int i = 1; i < m; ++i) {
            t = mul(t, t);
            r = mul(r, t);
        }
        return r;
    }
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
    // Barrett
    uint32_t p;
    unsigned long long mu;
    // lookup tables (m <= 16 built on host; staged in smem when m <= 12)
    uint16_t* d_exp;
    uint16_t* d_log;
    uint16_t* h_exp;
    uint16_t* h_log;
    // sticky error flag in mapped pinned host memory
    int* h_err;
    int* d_err;

    gfb::PrimeBarrett prime_barrett() const { return gfb::PrimeBarrett{p, mu}; }
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
        return gfb::Binary128{degree, pl, mask};
    }
};
