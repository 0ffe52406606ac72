// gfb_dot.cuh — inner-product accumulators shared by the matmul / convolution kernels (gfb_linalg.cu) and the
// blocked Gaussian elimination (gfb_solve.cu).  GF(p): unreduced 64-bit products summed in a 128-bit accumulator and
// reduced once per output; GF(2^m): xor of field products.
#pragma once
#include "gfb_common.cuh"

namespace gfb {

// ----------------------------------------------------------------------------------------
// dot-product accumulators: lazy for prime fields, xor for binary fields
// ----------------------------------------------------------------------------------------
template <typename F>
struct DotAcc {
    typename F::elem acc;
    __device__ __forceinline__ void init(const F& f) { acc = f.zero(); }
    __device__ __forceinline__ void fma(const F& f, typename F::elem a, typename F::elem b) { acc = f.add(acc, f.mul(a, b)); }
    __device__ __forceinline__ typename F::elem done(const F&) { return acc; }
};

template <>
struct DotAcc<PrimeBarrett> {
    unsigned long long lo, hi;
    __device__ __forceinline__ void init(const PrimeBarrett&) { lo = hi = 0ull; }
    __device__ __forceinline__ void fma(const PrimeBarrett&, uint32_t a, uint32_t b) {
        unsigned long long x = (unsigned long long)a * b;
        lo += x;
        hi += (lo < x) ? 1ull : 0ull;
    }
    __device__ __forceinline__ uint32_t done(const PrimeBarrett& f) {
        // (hi * 2^64 + lo) mod p, hi < 2^32
        uint32_t r64 = f.reduce64(0xFFFFFFFFFFFFFFFFull);  // (2^64 - 1) mod p
        r64 = f.add(r64, 1u % f.p);                        // 2^64 mod p
        uint32_t h = f.reduce64(hi);
        uint32_t l = f.reduce64(lo);
        return f.add(f.mul(h, r64), l);
    }
};

template <>
struct DotAcc<PrimeMersenne31> {
    unsigned long long lo, hi;
    __device__ __forceinline__ void init(const PrimeMersenne31&) { lo = hi = 0ull; }
    __device__ __forceinline__ void fma(const PrimeMersenne31&, uint32_t a, uint32_t b) {
        unsigned long long x = (unsigned long long)a * b;
        lo += x;
        hi += (lo < x) ? 1ull : 0ull;
    }
    __device__ __forceinline__ uint32_t done(const PrimeMersenne31& f) {
        // 2^64 = 2^(31*2+2) ≡ 4 (mod 2^31-1)
        const unsigned long long p = PrimeMersenne31::p;
        unsigned long long l = (lo & p) + ((lo >> 31) & p) + (lo >> 62);  // < 3 * 2^31
        unsigned long long h = (hi % p) * 4ull;                           // hi < 2^32
        unsigned long long s = (l + h) % p;
        return (uint32_t)s;
    }
};

}  // namespace gfb
