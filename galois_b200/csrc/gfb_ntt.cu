// gfb_ntt.cu — batched radix-2 NTT / INTT over prime fields for sm_100a.
//
// Replaces fft_jit / ifft_jit (/root/reference/src/galois/_domains/_function.py:170-392; radix-2
// butterfly :337-352) behind galois.ntt / galois.intt (_ntt.py:239-278).  The reference computes
//     X_k = sum_j x_j * omega^(j k)  (mod p),   natural order in and out,
// one 1-D vector per call.  Any exact algorithm yields the same X; only omega (chosen by the caller
// exactly as the reference does) and the optional 1/N scale matter for bit-exactness.
//
// Two device paths:
//  * four-step fast path, N = 2^20 = 1024 x 1024, p < 2^31 (the BASELINE.json config):
//      pass 1  : tile of 16 columns x 1024 rows → shared memory (transposed), one warp per column
//                runs a 1024-point NTT as two in-register radix-32 rounds with a shared-memory
//                exchange, multiplies by the four-step twiddle omega^(c*k) (x scale) and writes the
//                tile back in place-layout to a uint32 workspace.
//      pass 2  : tile of 16 rows x 1024 → shared memory, same warp NTT, transposed store
//                X[k1 + 1024*k2].
//    Butterflies are Harvey lazy butterflies with Shoup twiddles (values in [0,4p) when p < 2^30).
//    Work is issued in chunks of kNttChunk transforms (bounded uint32 workspace, whole waves of CTAs).
//  * generic path, any N = 2^k <= 2^24, any p < 2^32: one Stockham autosort radix-2 stage per
//    launch, ping-ponging between two workspaces (mirrors the reference's own stage structure).
//
// Algorithmic bytes per transform: 2 * N * elem_bytes (read x once, write X once).
#include <cstdlib>
#include <cstring>

#include "gfb_common.cuh"
#include "gfb_dispatch.cuh"
#include "gfb_host.cuh"

struct gfb_ntt_plan {
    gfb_field* field;
    size_t n;
    int log_n;
    uint32_t p;
    uint32_t omega, scale;
    bool fast;             // four-step n1 x 1024 path available (2^15 <= N <= 2^20, p < 2^31)
    bool lazy;             // p < 2^30
    // generic path
    uint32_t* d_pow;       // omega^e, e < max(n/2, 1)
    // fast path
    uint2* d_tw4;          // [k*1024 + c] = (scale * omega^(c k), shoup), k < n1 = N/1024
    uint2* d_zt;           // pass 2: [k1*32 + t] = (zeta2^(t k1), shoup), zeta2 = omega^n1 (1024th root), k1 < 32
    uint2* d_zt1;          // pass 1: [k1*32 + t] = (zeta1^(t k1), shoup), zeta1 = omega^1024 (n1-th root), k1 < n1/32
    uint32_t rho[16], rho_sh[16];  // rho^e, rho = zeta^32 (primitive 32nd root), with Shoup companions
    // mixed-radix path (any N | order - 1, prime or binary field): the reference's general stage, one launch per factor
    bool mixed;
    int n_factors;
    uint32_t factors[32];  // prime factors of N with multiplicity, ascending (fft_jit._prime_factors)
    uint64_t omega64, scale64;
    void* d_pow_mixed;     // omega^e, e < N, in the field's compute element type
    // hybrid N = n_odd * 2^k (k >= 5, p < 2^31): odd radices on columns, then the power-of-two kernels on rows
    gfb_ntt_plan* sub;     // plan of size n_pow2 with omega^n_odd (and the caller's scale)
    size_t n_odd, n_pow2;
    int n_odd_factors;     // the odd entries of factors[] (they follow the 2s)
};

namespace gfb {

constexpr int kNttChunk = 74;              // transforms per pass-1/pass-2 launch pair: 74*64 CTAs = 16 full waves of 2 CTAs/SM x 148 SMs
constexpr size_t kGenericWsBytes = (size_t)64 << 20;

// ------------------------------------------------------------------------------------------------
// table construction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t shoup(uint32_t w, uint32_t p) {
    return (uint32_t)((((unsigned long long)w) << 32) / p);
}

__global__ void k_build_pow(PrimeBarrett f, uint32_t omega, uint32_t* out, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = f.pow_u64(omega, i);
}

// tw4[k*1024 + c] = scale * omega^(c*k), k < n1
__global__ void k_build_tw4(PrimeBarrett f, uint32_t omega, uint32_t scale, uint2* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = i >> 10, c = i & 1023;
    uint32_t w = f.mul(f.pow_u64(omega, k * c), scale);
    out[i] = make_uint2(w, shoup(w, f.p));
}

// zt[k1*32 + t] = zeta^(t*k1), k1 < count/32
__global__ void k_build_zt(PrimeBarrett f, uint32_t zeta, uint2* out, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long k1 = i >> 5, t = i & 31;
    uint32_t w = f.pow_u64(zeta, k1 * t);
    out[i] = make_uint2(w, shoup(w, f.p));
}

// ------------------------------------------------------------------------------------------------
// generic path: one Stockham radix-2 stage per launch (structure of _function.py:337-352)
//   in_view[k, qi, b] = in[(k*q + qi)*m + b],  out_view[qi, f, b] = out[(qi*2 + f)*m + b],
//   twiddle(b) = omega^(q*b)
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_stockham_stage(PrimeBarrett f, const TI* __restrict__ in, TO* __restrict__ out,
                                                         const uint32_t* __restrict__ pow, size_t n, size_t m, size_t batch,
                                                         uint32_t scale) {
    const size_t half = n >> 1;
    const size_t total = half * batch;
    const size_t q = half / m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx / half;
        const size_t j = idx - row * half;  // j = qi*m + b
        const size_t qi = j / m, b = j - qi * m;
        const TI* src = in + row * n;
        TO* dst = out + row * n;
        uint32_t x0 = (uint32_t)src[qi * m + b];
        uint32_t x1 = f.mul((uint32_t)src[(q + qi) * m + b], pow[q * b]);
        uint32_t y0 = f.add(x0, x1), y1 = f.sub(x0, x1);
        if (scale != 1u) {
            y0 = f.mul(y0, scale);
            y1 = f.mul(y1, scale);
        }
        dst[(qi * 2) * m + b] = (TO)y0;
        dst[(qi * 2 + 1) * m + b] = (TO)y1;
    }
}

template <typename T>
__global__ void k_scale_copy(PrimeBarrett f, const T* in, T* out, size_t count, uint32_t scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = (T)f.mul((uint32_t)in[i], scale);
}

// ------------------------------------------------------------------------------------------------
// mixed-radix path: the general radix-r stage of fft_jit.implementation (_function.py:354-376), one output per thread
//   in_view[k, qi, b] = in[(k*q + qi)*m + b],  out_view[qi, fo, b] = out[(qi*r + fo)*m + b],
//   out = sum_k in_view[k] * tw^k (Horner),  tw = omega^(q*(fo*m + b))   — the reference's running twiddle
// Works over any field functor (prime or binary), any N dividing order - 1.
// ------------------------------------------------------------------------------------------------
template <typename F>
__global__ void k_build_pow_any(F f, typename F::elem omega, typename F::elem* out, size_t count, size_t fs) {
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = i == 0 ? f.one() : f.pow_u64(omega, i);
}

// ok[0] = 1 iff omega^n == 1 and omega^(n/q) != 1 for every prime factor q of n (omega has order exactly n)
template <typename F>
__global__ void k_check_root(F f, typename F::elem omega, unsigned long long n, const uint32_t* __restrict__ factors,
                             int nf, int* ok) {
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    bool good = !f.is_zero(omega) && f.pow_u64(omega, n) == f.one();
    for (int i = 0; i < nf && good; ++i) {
        if (i && factors[i] == factors[i - 1]) continue;
        if (f.pow_u64(omega, n / factors[i]) == f.one()) good = false;
    }
    *ok = good ? 1 : 0;
}

template <typename T, typename F>
__global__ void __launch_bounds__(256)
k_mixed_stage(F f, const T* __restrict__ in, T* __restrict__ out, const typename F::elem* __restrict__ pow, size_t n,
              size_t m, size_t r, size_t batch, typename F::elem scale, int apply_scale, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    const size_t total = n * batch;
    const size_t q = n / (m * r);
    const size_t rm = r * m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx / n;
        const size_t o = idx - row * n;
        const size_t qi = o / rm;
        const size_t rem = o - qi * rm;  // fo*m + b
        const size_t b = rem % m;
        const T* src = in + row * n + qi * m + b;
        const E tw = f.premul(pow[q * rem]);  // q * rem < q*r*m = n; multiplier form: one reduction per Horner step
        const size_t kstride = q * m;
        E acc = to_elem<E, T>(src[(r - 1) * kstride]);
        for (size_t k = r - 1; k-- > 0;) acc = f.add(f.mul_prepared(acc, tw), to_elem<E, T>(src[k * kstride]));
        if (apply_scale) acc = f.mul(acc, scale);
        out[idx] = to_store<T, E>(acc);
    }
}

// Hybrid form, N = N1 * V with V = 2^k: the same radix-r stage applied along axis 0 of each row viewed as [N1][V]
// (lane index v runs along the contiguous axis: coalesced), omega_N1 = omega^V so the stage twiddle is pow[V * q * rem].
// The last odd stage also applies the Cooley-Tukey twiddle omega^(k1 * v) between the two factors.
template <typename T, typename F>
__global__ void __launch_bounds__(256)
k_mixed_stage_v(F f, const T* __restrict__ in, T* __restrict__ out, const typename F::elem* __restrict__ pow, size_t n1,
                size_t V, size_t m, size_t r, size_t batch, int fuse_twiddle, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    const size_t n = n1 * V;
    const size_t total = n * batch;
    const size_t q = n1 / (m * r);
    const size_t rm = r * m;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx / n;
        const size_t within = idx - row * n;
        const size_t o = within / V;  // output index along axis 0 (k1 after the last stage)
        const size_t v = within - o * V;
        const size_t qi = o / rm;
        const size_t rem = o - qi * rm;  // fo*m + b
        const size_t b = rem % m;
        const T* src = in + row * n + (qi * m + b) * V + v;
        const E tw = f.premul(pow[q * rem * V]);  // multiplier form: one reduction per Horner step
        const size_t kstride = q * m * V;
        E acc = to_elem<E, T>(src[(r - 1) * kstride]);
        for (size_t k = r - 1; k-- > 0;) acc = f.add(f.mul_prepared(acc, tw), to_elem<E, T>(src[k * kstride]));
        if (fuse_twiddle) acc = f.mul(acc, pow[o * v]);  // o * v < N1 * V = N
        out[idx] = to_store<T, E>(acc);
    }
}

// out[row][k2 * n1 + k1] = in[row][k1 * V + k2]: the output permutation of the two-factor split, 32 x 32 tiles
template <typename T>
__global__ void __launch_bounds__(256) k_transpose_rows(const T* __restrict__ in, T* __restrict__ out, size_t n1, size_t V) {
    __shared__ T tile[32][33];
    const size_t row = blockIdx.z;
    const T* src = in + row * n1 * V;
    T* dst = out + row * n1 * V;
    const size_t k2_0 = (size_t)blockIdx.x * 32, k1_0 = (size_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const size_t k1 = k1_0 + j, k2 = k2_0 + tx;
        if (k1 < n1 && k2 < V) tile[j][tx] = src[k1 * V + k2];
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const size_t k2 = k2_0 + j, k1 = k1_0 + tx;
        if (k1 < n1 && k2 < V) dst[k2 * n1 + k1] = tile[tx][j];
    }
}

template <typename T, typename F>
__global__ void k_mixed_scale(F f, const T* __restrict__ in, T* __restrict__ out, size_t count, typename F::elem scale,
                              size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = to_store<T, E>(f.mul(to_elem<E, T>(in[i]), scale));
}

// ------------------------------------------------------------------------------------------------
// fast path arithmetic
// ------------------------------------------------------------------------------------------------
struct Radix32Roots {
    uint32_t w[16];
    uint32_t ws[16];
};

template <bool LAZY>
struct Mod {
    uint32_t p, p2;
    // conditional subtract without a compare: for x in [0, 2m), min(x, x - m) (the difference wraps when x < m)
    static __device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t m) { return min(x, x - m); }
    // x * w mod p with Shoup companion ws = floor(w 2^32 / p); any 32-bit x.  LAZY: result in [0, 2p).
    __device__ __forceinline__ uint32_t mul(uint32_t x, uint32_t w, uint32_t ws) const {
        uint32_t q = __umulhi(x, ws);
        uint32_t r = x * w - q * p;
        if constexpr (!LAZY) r = csub(r, p);
        return r;
    }
    // DIT butterfly (a, b) <- (a + t, a - t), t = b*w.  LAZY: in/out in [0, 4p); else in/out in [0, p).
    __device__ __forceinline__ void bfly(uint32_t& a, uint32_t& b, uint32_t w, uint32_t ws) const {
        if constexpr (LAZY) {
            const uint32_t x = csub(a, p2);
            const uint32_t t = mul(b, w, ws);
            a = x + t;
            b = x - t + p2;
        } else {
            const uint32_t t = mul(b, w, ws);
            const uint32_t s = csub(a + t, p);
            const uint32_t d = a - t;
            b = min(d, d + p);  // a < t: d wrapped, d + p is the residue
            a = s;
        }
    }
    // butterfly with unit twiddle
    __device__ __forceinline__ void bfly1(uint32_t& a, uint32_t& b) const {
        if constexpr (LAZY) {
            const uint32_t x = csub(a, p2), t = csub(b, p2);
            a = x + t;
            b = x - t + p2;
        } else {
            const uint32_t s = csub(a + b, p);
            const uint32_t d = a - b;
            b = min(d, d + p);
            a = s;
        }
    }
    // unit-twiddle butterfly whose inputs are known to be below 2p (canonical loads, Shoup products): no reductions
    __device__ __forceinline__ void bfly1_in2p(uint32_t& a, uint32_t& b) const {
        if constexpr (LAZY) {
            const uint32_t s = a + b;
            b = a - b + p2;
            a = s;
        } else {
            bfly1(a, b);
        }
    }
    // bring a lazy value into [0, p)
    __device__ __forceinline__ uint32_t canon(uint32_t x) const {
        if constexpr (LAZY) x = csub(x, p2);
        return csub(x, p);
    }
};

// In-register R-point DIT NTT, R = 2^LOG_R <= 32.  On entry v[brev(i)] = x[i]; on exit v[k] = sum_i x[i] rho_R^(i k)
// with rho_R = rho^(32/R): the twiddle of butterfly j in stage S is rho_R^(j R / 2^(S+1)) = rho^(j * (16 >> S)) for
// every R, so all sizes share the 16 radix-32 constants.
template <int LOG_R>
__host__ __device__ constexpr int brev(int i) {
    int r = 0;
    for (int b = 0; b < LOG_R; ++b) r |= ((i >> b) & 1) << (LOG_R - 1 - b);
    return r;
}
__host__ __device__ constexpr int brev5(int i) { return brev<5>(i); }

template <bool LAZY, int LOG_R, int S, bool IN2P = false>
__device__ __forceinline__ void radix_stage(uint32_t (&v)[1 << LOG_R], const Mod<LAZY>& md, const Radix32Roots& rt) {
    constexpr int half = 1 << S;
#pragma unroll
    for (int g = 0; g < (1 << LOG_R) / 2; ++g) {  // flat loop over the butterflies of stage S: every index is static
        const int j = g & (half - 1);
        const int k = (g >> S) << (S + 1);
        const int e = j * (16 >> S);
        if (e == 0) {
            if constexpr (IN2P)
                md.bfly1_in2p(v[k + j], v[k + j + half]);
            else
                md.bfly1(v[k + j], v[k + j + half]);
        } else {
            md.bfly(v[k + j], v[k + j + half], rt.w[e], rt.ws[e]);
        }
    }
}

// IN2P: every input is below 2p (stage 0 then needs no conditional subtractions in the lazy representation)
template <bool LAZY, int LOG_R, bool IN2P = false>
__device__ __forceinline__ void radixR(uint32_t (&v)[1 << LOG_R], const Mod<LAZY>& md, const Radix32Roots& rt) {
    if constexpr (LOG_R > 0) radix_stage<LAZY, LOG_R, 0, IN2P>(v, md, rt);
    if constexpr (LOG_R > 1) radix_stage<LAZY, LOG_R, 1>(v, md, rt);
    if constexpr (LOG_R > 2) radix_stage<LAZY, LOG_R, 2>(v, md, rt);
    if constexpr (LOG_R > 3) radix_stage<LAZY, LOG_R, 3>(v, md, rt);
    if constexpr (LOG_R > 4) radix_stage<LAZY, LOG_R, 4>(v, md, rt);
}

constexpr int kRowStride = 1058;  // uint32 words per shared-memory row (1024 + 32 pad + 2)
__device__ __forceinline__ int padi(int i) { return i + (i >> 5); }

// ---- n-point NTT (n = RA x 32, RA = 2^LOG_RA <= 32) of 16 vectors held in one shared-memory tile ------------------
// Thread mapping: w = tid & 15 selects the vector (a tile row), tq = tid >> 4 (0..15) and the sub-transform index
// s in {0,1} select t = tq + 16 s; the two half-warps therefore work on adjacent t.  Vector w lives at
// tile[w*kRowStride + padi(j)]: with kRowStride = 2 (mod 32) and one pad word per 32 elements, every access is
// bank-conflict free:
//   round A  (j = t + 32 a, lanes (w, t|t+1))  -> bank 2w + t + a     (even / odd halves)
//   round B  (j = b + 32 k1, lanes (w, k1|k1+1)) -> bank 2w + b + k1
//   row-wise (j = 32 consecutive, fixed w)     -> 32 consecutive words
// X[k1 + RA k0] = sum_{i0<32} rho^(i0 k0) zeta^(i0 k1) sum_{i1<RA} x[32 i1 + i0] rho_RA^(i1 k1),  zeta = omega_n.

constexpr int kFastThreads = 256;
constexpr int kTileW = 16;
constexpr size_t kFastSmem = (size_t)kTileW * kRowStride * sizeof(uint32_t);

// second half of round A: twiddle by zeta^(t k1) and park Y[k1] at j = t + 32 k1
template <bool LAZY, int LOG_RA>
__device__ __forceinline__ void round_a_store(uint32_t* row, int t, const uint32_t (&v)[1 << LOG_RA],
                                              const Mod<LAZY>& md, const uint2* __restrict__ zt) {
    row[t] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < (1 << LOG_RA); ++k1) {
        const uint2 z = __ldg(&zt[k1 * 32 + t]);
        row[t + 33 * k1] = md.mul(v[k1], z.x, z.y);
    }
}

// pass 1: column NTTs of length n1 = 32 RA + four-step twiddle.  One CTA per (transform, 16-column tile); N = n1 x 1024.
// Lanes run along the 16 columns, so the strided global reads (64 B segments) go straight into registers and the
// results go straight from registers to the workspace: the tile is only used for the exchange between the rounds.
template <typename T, bool LAZY, int LOG_RA>
__global__ void __launch_bounds__(kFastThreads, 2)
    k_ntt_pass1(Mod<LAZY> md, Radix32Roots rt, const uint2* __restrict__ zt, const uint2* __restrict__ tw4,
                const T* __restrict__ x, uint32_t* __restrict__ ws) {
    constexpr int RA = 1 << LOG_RA;
    constexpr int LOG_N = LOG_RA + 15;
    extern __shared__ uint32_t tile[];
    const size_t transform = blockIdx.x >> 6;
    const int c0 = (blockIdx.x & 63) * kTileW;
    const int w = threadIdx.x & 15, tq = threadIdx.x >> 4;
    const T* src = x + (transform << LOG_N) + c0 + w;
    uint32_t* dst = ws + (transform << LOG_N) + c0 + w;
    const uint2* twc = tw4 + c0 + w;
    uint32_t* row = tile + w * kRowStride;
    {
        // x[(32 a + t) * 1024 + c]: element i1 = a of the sub-transform i0 = t.  Both sub-transforms' loads are
        // issued before any arithmetic (up to 64 independent 4-byte loads per thread): one DRAM latency per CTA.
        uint32_t v[RA], u[RA];
#pragma unroll
        for (int a = 0; a < RA; ++a) v[brev<LOG_RA>(a)] = (uint32_t)__ldcs(src + (size_t)(32 * a + tq) * 1024);
#pragma unroll
        for (int a = 0; a < RA; ++a) u[brev<LOG_RA>(a)] = (uint32_t)__ldcs(src + (size_t)(32 * a + tq + 16) * 1024);
        radixR<LAZY, LOG_RA>(v, md, rt);
        round_a_store<LAZY, LOG_RA>(row, tq, v, md, zt);
        radixR<LAZY, LOG_RA>(u, md, rt);
        round_a_store<LAZY, LOG_RA>(row, tq + 16, u, md, zt);
    }
    __syncthreads();
    uint32_t v[32];
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int k1 = tq + 16 * s;
        if (k1 >= RA) break;  // n1 < 1024: fewer than 32 second-round transforms per vector
#pragma unroll
        for (int b = 0; b < 32; ++b) v[brev5(b)] = row[33 * k1 + b];
        radixR<LAZY, 5>(v, md, rt);
        // k = k1 + RA k0; multiply by scale * omega^(c k) and store ws[k][c]
#pragma unroll
        for (int k0 = 0; k0 < 32; ++k0) {
            const size_t k = (size_t)(k1 + RA * k0);
            const uint2 tw = __ldg(twc + k * 1024);
            uint32_t val = md.mul(v[k0], tw.x, tw.y);
            if constexpr (LAZY) val = Mod<LAZY>::csub(val, md.p);
            dst[k * 1024] = val;
        }
    }
}

// pass 2: the n1 row NTTs of length 1024 + transposed store.  One CTA per (transform, 16-row tile).  The 16 rows are
// contiguous in the workspace, so they are staged through the tile (coalesced along the row, transposed by the tile
// layout); the outputs X[k1 + n1 k2] have the 16 rows (k1) innermost and go straight from registers to global memory.
template <typename T, bool LAZY, bool SMALL, int LOG_N1_CT>
__global__ void __launch_bounds__(kFastThreads, 2)
    k_ntt_pass2(Mod<LAZY> md, Radix32Roots rt, const uint2* __restrict__ zt, const uint32_t* __restrict__ ws,
                T* __restrict__ y, int log_n1_rt, size_t rows_total) {
    extern __shared__ uint32_t tile[];
    // compile-time n1 for the large sizes keeps the 32 transposed stores on immediate offsets (no 64-bit IMADs)
    const int log_n1 = LOG_N1_CT >= 0 ? LOG_N1_CT : log_n1_rt;
    // the workspace is [transform][k1 < n1][1024]: rows are contiguous across transforms, a tile is 16 consecutive rows
    const size_t row0 = (size_t)blockIdx.x * kTileW;
    const size_t n1 = (size_t)1 << log_n1;
    const uint32_t* src = ws + row0 * 1024;
    const int w = threadIdx.x & 15, tq = threadIdx.x >> 4;
    const size_t my_row = row0 + w;
    // output X[transform][k1 + n1 k2]; with n1 >= 16 the 16 rows of a tile are 16 consecutive k1 of one transform
    T* dst = y + ((my_row >> log_n1) << (log_n1 + 10)) + (my_row & (n1 - 1));
    const bool row_live = !SMALL || my_row < rows_total;
    // 16 rows x 1024 contiguous words = 4096 uint4: 16 per thread, consecutive threads consecutive uint4
    {
        const uint4* src4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
        for (int it = 0; it < 16; ++it) {
            const int e = it * kFastThreads + threadIdx.x;  // uint4 index within the tile
            const int ww = e >> 8, i = (e & 255) * 4;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (!SMALL || row0 + ww < rows_total) q = __ldg(src4 + e);
            uint32_t* d = tile + ww * kRowStride + padi(i);  // i % 4 == 0: the four words share one pad offset
            d[0] = q.x;
            d[1] = q.y;
            d[2] = q.z;
            d[3] = q.w;
        }
    }
    __syncthreads();
    uint32_t* row = tile + w * kRowStride;
    uint32_t v[32];
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int t = tq + 16 * s;
#pragma unroll
        for (int a = 0; a < 32; ++a) v[brev5(a)] = row[t + 33 * a];
        radixR<LAZY, 5>(v, md, rt);
        round_a_store<LAZY, 5>(row, t, v, md, zt);  // in place: thread (w, t) owns positions {t + 32 a}
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const int t = tq + 16 * s;  // k1 of the inner transform
#pragma unroll
        for (int b = 0; b < 32; ++b) v[brev5(b)] = row[33 * t + b];
        radixR<LAZY, 5>(v, md, rt);
        if (row_live) {
#pragma unroll
            for (int k0 = 0; k0 < 32; ++k0) dst[(size_t)(t + 32 * k0) * n1] = (T)md.canon(v[k0]);
        }
    }
}

// pass 1 for short columns, n1 = 2^LOG_N1 in {2, 4, 8, 16} (N = 2^11 .. 2^14): one thread per column does the whole
// n1-point transform in registers; lanes run along the row, so every access is a full 128-byte line.
template <typename T, bool LAZY, int LOG_N1>
__global__ void __launch_bounds__(256)
    k_ntt_pass1_small(Mod<LAZY> md, Radix32Roots rt, const uint2* __restrict__ tw4, const T* __restrict__ x,
                      uint32_t* __restrict__ ws, size_t columns_total) {
    constexpr int N1 = 1 << LOG_N1;
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // transform * 1024 + c
    if (col >= columns_total) return;
    const size_t transform = col >> 10, c = col & 1023;
    const T* src = x + (transform << (LOG_N1 + 10)) + c;
    uint32_t* dst = ws + (transform << (LOG_N1 + 10)) + c;
    static_assert(LOG_N1 <= 5, "one in-register radix stage set covers at most 32 points");
    uint32_t v[N1];
#pragma unroll
    for (int r = 0; r < N1; ++r) v[brev<LOG_N1>(r)] = (uint32_t)__ldcs(src + (size_t)r * 1024);
    radixR<LAZY, LOG_N1>(v, md, rt);
#pragma unroll
    for (int k = 0; k < N1; ++k) {
        const uint2 tw = __ldg(tw4 + (size_t)k * 1024 + c);
        uint32_t val = md.mul(v[k], tw.x, tw.y);
        if constexpr (LAZY) val = Mod<LAZY>::csub(val, md.p);
        dst[(size_t)k * 1024] = val;
    }
}

// pass 1 for n1 = 64 (N = 2^16): still one thread per column — a radix-2 step across the two 32-row halves, the
// zeta1^(i0) twiddle, then two in-register radix-32 transforms (X[k1 + 2 k0], k1 in {0,1}).
template <typename T, bool LAZY>
__global__ void __launch_bounds__(256)
    k_ntt_pass1_c64(Mod<LAZY> md, Radix32Roots rt, const uint2* __restrict__ zt1, const uint2* __restrict__ tw4,
                    const T* __restrict__ x, uint32_t* __restrict__ ws, size_t columns_total) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= columns_total) return;
    const size_t transform = col >> 10, c = col & 1023;
    const T* src = x + (transform << 16) + c;
    uint32_t* dst = ws + (transform << 16) + c;
    uint32_t a0[32], a1[32];
#pragma unroll
    for (int i0 = 0; i0 < 32; ++i0) {
        a0[brev5(i0)] = (uint32_t)__ldcs(src + (size_t)i0 * 1024);
        a1[brev5(i0)] = (uint32_t)__ldcs(src + (size_t)(32 + i0) * 1024);
    }
#pragma unroll
    for (int i0 = 0; i0 < 32; ++i0) {
        md.bfly1(a0[brev5(i0)], a1[brev5(i0)]);  // Y[0][i0], Y[1][i0]
        if (i0 > 0) {
            const uint2 z = __ldg(&zt1[32 + i0]);  // zeta1^(i0 * 1)
            a1[brev5(i0)] = md.mul(a1[brev5(i0)], z.x, z.y);
        }
    }
    radixR<LAZY, 5>(a0, md, rt);
    radixR<LAZY, 5>(a1, md, rt);
#pragma unroll
    for (int k0 = 0; k0 < 32; ++k0) {
#pragma unroll
        for (int k1 = 0; k1 < 2; ++k1) {
            const size_t k = (size_t)(k1 + 2 * k0);
            const uint2 tw = __ldg(tw4 + k * 1024 + c);
            uint32_t val = md.mul(k1 ? a1[k0] : a0[k0], tw.x, tw.y);
            if constexpr (LAZY) val = Mod<LAZY>::csub(val, md.p);
            dst[k * 1024] = val;
        }
    }
}

// whole transforms of N = 32 RA <= 1024 points, 16 per CTA, in one kernel: coalesced staged load, the two rounds,
// results parked in the tile in natural order, coalesced store (with the optional 1/N scale).
template <typename T, bool LAZY, int LOG_RA>
__global__ void __launch_bounds__(kFastThreads, 2)
    k_ntt_small(Mod<LAZY> md, Radix32Roots rt, const uint2* __restrict__ zt, const T* __restrict__ x, T* __restrict__ y,
                size_t batch, uint32_t scale, uint32_t scale_sh) {
    constexpr int RA = 1 << LOG_RA;
    constexpr int LOG_N = LOG_RA + 5;
    constexpr int N = 1 << LOG_N;
    extern __shared__ uint32_t tile[];
    const size_t row0 = (size_t)blockIdx.x * kTileW;
    const size_t rows_here = batch - row0 < (size_t)kTileW ? batch - row0 : (size_t)kTileW;
    const T* src = x + row0 * N;
    T* dst = y + row0 * N;
    const int w = threadIdx.x & 15, tq = threadIdx.x >> 4;
    for (int e = threadIdx.x; e < kTileW * N; e += kFastThreads) {
        const int ww = e >> LOG_N, i = e & (N - 1);
        tile[ww * kRowStride + padi(i)] = (size_t)ww < rows_here ? (uint32_t)src[e] : 0u;
    }
    __syncthreads();
    uint32_t* row = tile + w * kRowStride;
    {
        uint32_t v[RA];
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            const int t = tq + 16 * s;
#pragma unroll
            for (int a = 0; a < RA; ++a) v[brev<LOG_RA>(a)] = row[t + 33 * a];
            radixR<LAZY, LOG_RA>(v, md, rt);
            round_a_store<LAZY, LOG_RA>(row, t, v, md, zt);
        }
    }
    __syncthreads();
    // round B reads rows {33 k1 + b} and writes positions k1 + RA k0 that other threads still read: keep both
    // sub-transforms in registers until every thread has finished reading
    uint32_t v0[32], v1[32];
    const bool has0 = tq < RA, has1 = tq + 16 < RA;
    if (has0) {
#pragma unroll
        for (int b = 0; b < 32; ++b) v0[brev5(b)] = row[33 * tq + b];
    }
    if (has1) {
#pragma unroll
        for (int b = 0; b < 32; ++b) v1[brev5(b)] = row[33 * (tq + 16) + b];
    }
    __syncthreads();
    if (has0) {
        radixR<LAZY, 5>(v0, md, rt);
#pragma unroll
        for (int k0 = 0; k0 < 32; ++k0) row[padi(tq + RA * k0)] = v0[k0];
    }
    if (has1) {
        radixR<LAZY, 5>(v1, md, rt);
#pragma unroll
        for (int k0 = 0; k0 < 32; ++k0) row[padi(tq + 16 + RA * k0)] = v1[k0];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kTileW * N; e += kFastThreads) {
        const int ww = e >> LOG_N, i = e & (N - 1);
        if ((size_t)ww < rows_here) {
            uint32_t val = tile[ww * kRowStride + padi(i)];
            if (scale != 1u) {
                val = md.mul(val, scale, scale_sh);
                if constexpr (LAZY) val = Mod<LAZY>::csub(val, md.p);
            } else {
                val = md.canon(val);
            }
            dst[e] = (T)val;
        }
    }
}

#include "gfb_ntt_fused.cuh"

// transforms per pass-1 / pass-2 launch pair: 74 * 2^20 elements of uint32 workspace; at N = 2^20 that is
// 74 * 64 CTAs = 16 full waves of 2 CTAs/SM x 148 SMs
// Tuning knob (experiments only): GFB_NTT_CHUNK=<transforms of 2^20> overrides the chunk, e.g. 18 keeps the 72 MiB
// workspace of a chunk inside the 126 MB L2 between the two passes at the price of a ragged last wave.
static size_t fast_chunk(const gfb_ntt_plan* pl) {
    static const int base = [] {
        const char* e = getenv("GFB_NTT_CHUNK");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 1024 ? v : kNttChunk;
    }();
    return (size_t)base << (20 - pl->log_n);
}

static Radix32Roots roots_of(const gfb_ntt_plan* pl) {
    Radix32Roots rt;
    memcpy(rt.w, pl->rho, sizeof(rt.w));
    memcpy(rt.ws, pl->rho_sh, sizeof(rt.ws));
    return rt;
}

// N = n1 x 1024, n1 = 32 RA (2^15 .. 2^20)
template <typename T, bool LAZY, int LOG_RA>
static gfb_status run_fast_ra(gfb_ntt_plan* pl, const T* x, T* y, size_t batch, uint32_t* ws, cudaStream_t st) {
    Mod<LAZY> md{pl->p, pl->p * 2u};
    const Radix32Roots rt = roots_of(pl);
    // per-device attribute (several devices per process are supported): set on every call, it is cheap
    GFB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass1<T, LAZY, LOG_RA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kFastSmem));
    GFB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass2<T, LAZY, false, LOG_RA + 5>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
    const int log_n1 = LOG_RA + 5;
    const size_t chunk = fast_chunk(pl);
    for (size_t off = 0; off < batch; off += chunk) {
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        const T* xin = x + (off << pl->log_n);
        T* yout = y + (off << pl->log_n);
        k_ntt_pass1<T, LAZY, LOG_RA>
            <<<(unsigned)(cnt * 64), kFastThreads, kFastSmem, st>>>(md, rt, pl->d_zt1, pl->d_tw4, xin, ws);
        k_ntt_pass2<T, LAZY, false, LOG_RA + 5><<<(unsigned)(cnt << (log_n1 - 4)), kFastThreads, kFastSmem, st>>>(
            md, rt, pl->d_zt, ws, yout, log_n1, cnt << log_n1);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

// N = n1 x 1024, n1 = 2 .. 64 (2^11 .. 2^16): columns short enough for one thread each
template <typename T, bool LAZY, int LOG_N1>
static gfb_status run_fast_short(gfb_ntt_plan* pl, const T* x, T* y, size_t batch, uint32_t* ws, cudaStream_t st) {
    Mod<LAZY> md{pl->p, pl->p * 2u};
    const Radix32Roots rt = roots_of(pl);
    // per-device attribute (several devices per process are supported): set on every call, it is cheap
    GFB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_pass2<T, LAZY, (LOG_N1 < 4), LOG_N1>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFastSmem));
    const size_t chunk = fast_chunk(pl);
    for (size_t off = 0; off < batch; off += chunk) {
        const size_t cnt = batch - off < chunk ? batch - off : chunk;
        const T* xin = x + (off << pl->log_n);
        T* yout = y + (off << pl->log_n);
        const size_t columns = cnt << 10, rows = cnt << LOG_N1;
        if constexpr (LOG_N1 == 6)
            k_ntt_pass1_c64<T, LAZY><<<(unsigned)((columns + 255) / 256), 256, 0, st>>>(md, rt, pl->d_zt1, pl->d_tw4, xin,
                                                                                        ws, columns);
        else
            k_ntt_pass1_small<T, LAZY, LOG_N1><<<(unsigned)((columns + 255) / 256), 256, 0, st>>>(md, rt, pl->d_tw4, xin,
                                                                                                   ws, columns);
        k_ntt_pass2<T, LAZY, (LOG_N1 < 4), LOG_N1>
            <<<(unsigned)((rows + kTileW - 1) / kTileW), kFastThreads, kFastSmem, st>>>(md, rt, pl->d_zt, ws, yout, LOG_N1,
                                                                                        rows);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

// N = 32 RA <= 1024: one kernel
template <typename T, bool LAZY, int LOG_RA>
static gfb_status run_small_ra(gfb_ntt_plan* pl, const T* x, T* y, size_t batch, cudaStream_t st) {
    Mod<LAZY> md{pl->p, pl->p * 2u};
    const Radix32Roots rt = roots_of(pl);
    // per-device attribute (several devices per process are supported): set on every call, it is cheap
    GFB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_small<T, LAZY, LOG_RA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kFastSmem));
    const uint32_t scale_sh = (uint32_t)((((unsigned long long)pl->scale) << 32) / pl->p);
    k_ntt_small<T, LAZY, LOG_RA><<<(unsigned)((batch + kTileW - 1) / kTileW), kFastThreads, kFastSmem, st>>>(
        md, rt, pl->d_zt, x, y, batch, pl->scale, scale_sh);
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

// N = 2^20, uint32 elements: the persistent fused kernel (gfb_ntt_fused.cuh).  GFB_NTT_FUSED=0 selects the two-launch
// kernels above instead (A/B measurements).
static bool fused_enabled() {
    static const bool on = [] {
        const char* e = getenv("GFB_NTT_FUSED");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <bool LAZY, int W>
static gfb_status run_fused_w(gfb_ntt_plan* pl, const uint32_t* x, uint32_t* y, size_t batch, uint32_t* ws, cudaStream_t st) {
    using namespace fused;
    using C = Cfg<W>;
    Mod<LAZY> md{pl->p, pl->p * 2u};
    const Radix32Roots rt = roots_of(pl);
    const size_t kMaxBatch = (size_t)1 << 23;  // tickets are 32-bit: 2 * 1024 / W per transform
    static const unsigned lag_cfg = [] {
        const char* e = getenv("GFB_NTT_LAG");
        const int v = e ? atoi(e) : 0;
        return (unsigned)(v >= 1 && v < (int)kRingMax ? v : (int)kLagDefault);
    }();
    static const unsigned ring_cfg = [] {
        const char* e = getenv("GFB_NTT_RING");
        const int v = e ? atoi(e) : 0;
        return (unsigned)(v >= 2 && v <= (int)kRingMax ? v : (int)kRingDefault);
    }();
    const unsigned ring_want = ring_cfg > lag_cfg ? ring_cfg : lag_cfg + 1;  // a slot is recycled only after its pass 2
    for (size_t off = 0; off < batch; off += kMaxBatch) {
        const size_t cnt = batch - off < kMaxBatch ? batch - off : kMaxBatch;
        const unsigned ring = (unsigned)(cnt < ring_want ? cnt : ring_want);
        Ctrl* ctrl = reinterpret_cast<Ctrl*>(ws + ((size_t)ring << 20));
        GFB_CUDA_TRY(cudaMemsetAsync(ctrl, 0, sizeof(Ctrl) + 2 * cnt * sizeof(unsigned int), st));
        GFB_CUDA_TRY(cudaFuncSetAttribute(k_ntt_fused<LAZY, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)C::kSmemBytes));
        size_t grid = (size_t)sm_count() * (W == 16 ? 1 : 2);
        if (grid > cnt * 2 * C::kTilesPerPass) grid = cnt * 2 * C::kTilesPerPass;
        k_ntt_fused<LAZY, W><<<(unsigned)grid, C::kThreads, C::kSmemBytes, st>>>(
            md, rt, pl->d_zt, pl->d_tw4, x + (off << 20), y + (off << 20), ws, ctrl, (unsigned)cnt, ring, lag_cfg);
        GFB_CUDA_TRY(cudaGetLastError());
    }
    return GFB_OK;
}

// tile width: 16 transforms per tile and one CTA per SM, or 8 and two (GFB_NTT_TILE)
template <bool LAZY>
static gfb_status run_fused(gfb_ntt_plan* pl, const uint32_t* x, uint32_t* y, size_t batch, uint32_t* ws, cudaStream_t st) {
    static const int tile_w = [] {
        const char* e = getenv("GFB_NTT_TILE");
        const int v = e ? atoi(e) : 0;
        return v == 8 || v == 16 ? v : 16;
    }();
    if (tile_w == 8) return run_fused_w<LAZY, 8>(pl, x, y, batch, ws, st);
    return run_fused_w<LAZY, 16>(pl, x, y, batch, ws, st);
}

template <typename T, bool LAZY>
static gfb_status run_fast_t(gfb_ntt_plan* pl, const T* x, T* y, size_t batch, uint32_t* ws, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        if (pl->log_n == 20 && fused_enabled() && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)ws) & 15) == 0)
            return run_fused<LAZY>(pl, (const uint32_t*)x, (uint32_t*)y, batch, ws, st);
    }
    switch (pl->log_n) {
        case 5: return run_small_ra<T, LAZY, 0>(pl, x, y, batch, st);
        case 6: return run_small_ra<T, LAZY, 1>(pl, x, y, batch, st);
        case 7: return run_small_ra<T, LAZY, 2>(pl, x, y, batch, st);
        case 8: return run_small_ra<T, LAZY, 3>(pl, x, y, batch, st);
        case 9: return run_small_ra<T, LAZY, 4>(pl, x, y, batch, st);
        case 10: return run_small_ra<T, LAZY, 5>(pl, x, y, batch, st);
        case 11: return run_fast_short<T, LAZY, 1>(pl, x, y, batch, ws, st);
        case 12: return run_fast_short<T, LAZY, 2>(pl, x, y, batch, ws, st);
        case 13: return run_fast_short<T, LAZY, 3>(pl, x, y, batch, ws, st);
        case 14: return run_fast_short<T, LAZY, 4>(pl, x, y, batch, ws, st);
        case 15: return run_fast_short<T, LAZY, 5>(pl, x, y, batch, ws, st);
        case 16: return run_fast_short<T, LAZY, 6>(pl, x, y, batch, ws, st);
        case 17: return run_fast_ra<T, LAZY, 2>(pl, x, y, batch, ws, st);
        case 18: return run_fast_ra<T, LAZY, 3>(pl, x, y, batch, ws, st);
        case 19: return run_fast_ra<T, LAZY, 4>(pl, x, y, batch, ws, st);
        case 20: return run_fast_ra<T, LAZY, 5>(pl, x, y, batch, ws, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename T>
static gfb_status run_generic_t(gfb_ntt_plan* pl, const T* x, T* y, size_t batch, void* workspace, cudaStream_t st) {
    PrimeBarrett f = pl->field->prime_barrett();
    const size_t n = pl->n;
    if (n == 1) {
        const size_t cnt = batch;
        k_scale_copy<T><<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(f, x, y, cnt, pl->scale);
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    }
    const size_t rows_per_chunk = kGenericWsBytes / (n * sizeof(uint32_t)) ? kGenericWsBytes / (n * sizeof(uint32_t)) : 1;
    uint32_t* A = (uint32_t*)workspace;
    uint32_t* B = A + rows_per_chunk * n;
    const int S = pl->log_n;
    for (size_t off = 0; off < batch; off += rows_per_chunk) {
        const size_t cnt = batch - off < rows_per_chunk ? batch - off : rows_per_chunk;
        const size_t total = (n >> 1) * cnt;
        size_t blocks = (total + 255) / 256;
        const size_t cap = (size_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        const unsigned grid = (unsigned)blocks;
        const T* xin = x + off * n;
        T* yout = y + off * n;
        size_t m = 1;
        for (int s = 0; s < S; ++s, m <<= 1) {
            const bool first = s == 0, last = s == S - 1;
            const uint32_t sc = last ? pl->scale : 1u;
            uint32_t* src_ws = (s & 1) ? A : B;  // stage s reads what stage s-1 wrote
            uint32_t* dst_ws = (s & 1) ? B : A;
            if (first && last)
                k_stockham_stage<T, T><<<grid, 256, 0, st>>>(f, xin, yout, pl->d_pow, n, m, cnt, sc);
            else if (first)
                k_stockham_stage<T, uint32_t><<<grid, 256, 0, st>>>(f, xin, dst_ws, pl->d_pow, n, m, cnt, sc);
            else if (last)
                k_stockham_stage<uint32_t, T><<<grid, 256, 0, st>>>(f, src_ws, yout, pl->d_pow, n, m, cnt, sc);
            else
                k_stockham_stage<uint32_t, uint32_t><<<grid, 256, 0, st>>>(f, src_ws, dst_ws, pl->d_pow, n, m, cnt, sc);
        }
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

static size_t mixed_rows_per_chunk(const gfb_ntt_plan* pl, int eb) {
    size_t rows = kGenericWsBytes / (pl->n * (size_t)eb);
    if (rows > 32768) rows = 32768;  // the row index is a grid dimension in the output permutation
    return rows ? rows : 1;
}

extern "C" gfb_status gfb_ntt_exec(gfb_ntt_plan* pl, int eb, const void* x, void* y, size_t batch, void* workspace,
                                   void* stream);
extern "C" size_t gfb_ntt_workspace_bytes(const gfb_ntt_plan* pl, int elem_bytes, size_t batch);

// N = n_odd * 2^k: odd radices down the columns (+ inter-factor twiddle), the power-of-two kernels along the rows,
// then the [n_odd][2^k] -> [2^k][n_odd] output permutation.  Workspace: two row buffers + the sub-plan's scratch.
template <typename T, typename F>
static gfb_status run_hybrid_t(gfb_ntt_plan* pl, F f, const T* x, T* y, size_t batch, void* workspace, cudaStream_t st) {
    using E = typename F::elem;
    const size_t n = pl->n, n1 = pl->n_odd, V = pl->n_pow2;
    const size_t fs = field_smem_bytes(f);
    const E* pow = (const E*)pl->d_pow_mixed;
    const size_t rows_per_chunk = mixed_rows_per_chunk(pl, (int)sizeof(T));
    const size_t rows_ws = batch < rows_per_chunk ? batch : rows_per_chunk;
    T* A = (T*)workspace;
    T* B = A + rows_ws * n;
    void* sub_ws = (void*)(B + rows_ws * n);
    const int S = pl->n_odd_factors;
    for (size_t off = 0; off < batch; off += rows_per_chunk) {
        const size_t cnt = batch - off < rows_per_chunk ? batch - off : rows_per_chunk;
        size_t blocks = (n * cnt + 255) / 256;
        const size_t cap = (size_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        const T* src = x + off * n;
        T* bufs[2] = {A, B};
        int cur = 0;
        size_t m = 1;
        for (int s = 0; s < S; ++s) {
            const size_t r = pl->factors[pl->n_factors - 1 - s];  // the odd radices sit at the end: largest first
            T* dst = bufs[cur];
            k_mixed_stage_v<T, F><<<(unsigned)blocks, 256, fs, st>>>(f, src, dst, pow, n1, V, m, r, cnt, s == S - 1, fs);
            src = dst;
            cur ^= 1;
            m *= r;
        }
        // rows of length V through the power-of-two plan (cnt * n1 independent transforms), into the other buffer
        T* rows_out = bufs[cur];
        gfb_status stt = gfb_ntt_exec(pl->sub, (int)sizeof(T), src, rows_out, cnt * n1, sub_ws, (void*)st);
        if (stt != GFB_OK) return stt;
        dim3 grid((unsigned)((V + 31) / 32), (unsigned)((n1 + 31) / 32), (unsigned)cnt);
        if (cnt > 65535 || grid.y > 65535) return GFB_ERR_UNSUPPORTED;
        k_transpose_rows<T><<<grid, 256, 0, st>>>(rows_out, y + off * n, n1, V);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename T, typename F>
static gfb_status run_mixed_t(gfb_ntt_plan* pl, F f, const T* x, T* y, size_t batch, void* workspace, cudaStream_t st) {
    using E = typename F::elem;
    const size_t n = pl->n;
    const size_t fs = field_smem_bytes(f);
    const E scale = (E)pl->scale64;
    const E* pow = (const E*)pl->d_pow_mixed;
    if (pl->sub) return run_hybrid_t<T, F>(pl, f, x, y, batch, workspace, st);
    const int S = pl->n_factors;
    if (S == 0) {  // N = 1
        k_mixed_scale<T, F><<<(unsigned)((batch + 255) / 256), 256, fs, st>>>(f, x, y, batch, scale, fs);
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    }
    const size_t rows_per_chunk = mixed_rows_per_chunk(pl, (int)sizeof(T));
    // the two ping-pong buffers are laid out as gfb_ntt_workspace_bytes sized them: min(batch, chunk) rows each
    const size_t rows_ws = batch < rows_per_chunk ? batch : rows_per_chunk;
    T* A = (T*)workspace;
    T* B = A ? A + rows_ws * n : nullptr;
    for (size_t off = 0; off < batch; off += rows_per_chunk) {
        const size_t cnt = batch - off < rows_per_chunk ? batch - off : rows_per_chunk;
        size_t blocks = (n * cnt + 255) / 256;
        const size_t cap = (size_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        const T* src = x + off * n;
        size_t m = 1;
        for (int s = 0; s < S; ++s) {
            const size_t r = pl->factors[S - 1 - s];  // radices are consumed from the end (_function.py:312-314)
            const bool last = s == S - 1;
            T* dst = last ? y + off * n : ((s & 1) ? B : A);
            k_mixed_stage<T, F><<<(unsigned)blocks, 256, fs, st>>>(f, src, dst, pow, n, m, r, cnt, scale,
                                                                     last && pl->scale64 != 1ull, fs);
            src = dst;
            m *= r;
        }
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

static gfb_status run_mixed(gfb_ntt_plan* pl, int eb, const void* x, void* y, size_t batch, void* workspace,
                            cudaStream_t st) {
    auto go = [&](auto functor) -> gfb_status {
        if (eb == 4) return run_mixed_t<uint32_t>(pl, functor, (const uint32_t*)x, (uint32_t*)y, batch, workspace, st);
        return run_mixed_t<unsigned long long>(pl, functor, (const unsigned long long*)x, (unsigned long long*)y, batch,
                                               workspace, st);
    };
    const gfb_field* f = pl->field;
    switch (f->impl) {
        case IMPL_PRIME_M31: return go(f->prime_m31());
        case IMPL_PRIME_BARRETT: return go(f->prime_barrett());
        case IMPL_BIN_LUT: return go(f->bin_lut());
        case IMPL_BIN_CALC: return go(f->bin_calc());
        default: break;
    }
    return GFB_ERR_UNSUPPORTED;
}

// tables + primitivity check of the mixed-radix plan
template <typename F>
static cudaError_t mixed_setup(gfb_ntt_plan* pl, F f, int* ok_out) {
    using E = typename F::elem;
    const size_t n = pl->n;
    const size_t fs = field_smem_bytes(f);
    uint32_t* d_factors = nullptr;
    int* d_ok = nullptr;
    cudaError_t e = cudaMalloc((void**)&pl->d_pow_mixed, n * sizeof(E));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_factors, sizeof(pl->factors));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_ok, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(d_factors, pl->factors, sizeof(pl->factors), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        k_check_root<F><<<1, 32, fs>>>(f, (E)pl->omega64, (unsigned long long)n, d_factors, pl->n_factors, d_ok);
        k_build_pow_any<F><<<(unsigned)((n + 255) / 256), 256, fs>>>(f, (E)pl->omega64, (E*)pl->d_pow_mixed, n, fs);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(ok_out, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(d_factors);
    cudaFree(d_ok);
    return e;
}

static uint32_t host_powmod(uint32_t b, unsigned long long e, uint32_t p) {
    unsigned long long r = 1 % p, x = b % p;
    while (e) {
        if (e & 1) r = r * x % p;  // p < 2^32: products fit 64 bits
        x = x * x % p;
        e >>= 1;
    }
    return (uint32_t)r;
}

}  // namespace gfb

extern "C" {

gfb_status gfb_ntt_plan_create(gfb_ntt_plan** out, gfb_field* f, size_t n, uint64_t omega, uint64_t scale) {
    using namespace gfb;
    if (!out) return GFB_ERR_BAD_ARG;
    *out = nullptr;
    if (!f || n == 0) return GFB_ERR_BAD_ARG;
    if (n > ((size_t)1 << 24)) return GFB_ERR_UNSUPPORTED;
    if (f->kind != GFB_PRIME || (n & (n - 1)) != 0) {
        // any N | order - 1 over a prime or binary field: the reference's mixed-radix stages
        if (f->impl == IMPL_BIN_128 || f->order_lo == 0) return GFB_ERR_UNSUPPORTED;
        if ((f->order_lo - 1) % n != 0) return GFB_ERR_BAD_ARG;
        if (omega >= f->order_lo || scale >= f->order_lo || scale == 0) return GFB_ERR_BAD_ARG;
        gfb_ntt_plan* pl = (gfb_ntt_plan*)calloc(1, sizeof(gfb_ntt_plan));
        pl->field = f;
        pl->n = n;
        pl->mixed = true;
        pl->omega64 = omega;
        pl->scale64 = scale;
        size_t rest = n;
        for (size_t d = 2; d * d <= rest; ++d)
            while (rest % d == 0) {
                pl->factors[pl->n_factors++] = (uint32_t)d;
                rest /= d;
            }
        if (rest > 1) pl->factors[pl->n_factors++] = (uint32_t)rest;
        int ok = 0;
        cudaError_t e = cudaSuccess;
        switch (f->impl) {
            case IMPL_PRIME_M31: e = mixed_setup(pl, f->prime_m31(), &ok); break;
            case IMPL_PRIME_BARRETT: e = mixed_setup(pl, f->prime_barrett(), &ok); break;
            case IMPL_BIN_LUT: e = mixed_setup(pl, f->bin_lut(), &ok); break;
            default: e = mixed_setup(pl, f->bin_calc(), &ok); break;
        }
        if (e != cudaSuccess) {
            set_last_error("gfb_ntt_plan_create (mixed radix)", e);
            gfb_ntt_plan_destroy(pl);
            return GFB_ERR_CUDA;
        }
        if (!ok) {  // omega is not a primitive n-th root of unity
            gfb_ntt_plan_destroy(pl);
            return GFB_ERR_BAD_ARG;
        }
        // N = n_odd * 2^k with 2^k a size the power-of-two kernels take (32 .. 2^20, p < 2^31): hybrid split
        size_t pow2 = 1;
        int twos = 0;
        while (twos < pl->n_factors && pl->factors[twos] == 2) {
            pow2 <<= 1;
            ++twos;
        }
        if (f->kind == GFB_PRIME && f->p < (1u << 31) && twos >= 5 && twos <= 20 && twos < pl->n_factors) {
            pl->n_pow2 = pow2;
            pl->n_odd = n / pow2;
            pl->n_odd_factors = pl->n_factors - twos;
            const uint32_t omega_sub = host_powmod((uint32_t)omega, pl->n_odd, f->p);  // order 2^k
            gfb_status ss = gfb_ntt_plan_create(&pl->sub, f, pow2, omega_sub, scale);
            if (ss != GFB_OK) {
                gfb_ntt_plan_destroy(pl);
                return ss;
            }
        }
        *out = pl;
        return GFB_OK;
    }
    const uint32_t p = f->p;
    if (omega >= p || scale >= p || scale == 0) return GFB_ERR_BAD_ARG;
    if (((unsigned long long)p - 1ull) % n != 0) return GFB_ERR_BAD_ARG;
    // omega must be a primitive n-th root of unity
    if (host_powmod((uint32_t)omega, n, p) != 1u % p) return GFB_ERR_BAD_ARG;
    if (n > 1 && host_powmod((uint32_t)omega, n / 2, p) == 1u) return GFB_ERR_BAD_ARG;

    gfb_ntt_plan* pl = (gfb_ntt_plan*)calloc(1, sizeof(gfb_ntt_plan));
    pl->field = f;
    pl->n = n;
    pl->log_n = 0;
    while (((size_t)1 << pl->log_n) < n) pl->log_n++;
    pl->p = p;
    pl->omega = (uint32_t)omega;
    pl->scale = (uint32_t)scale;
    pl->fast = pl->log_n >= 5 && pl->log_n <= 20 && p < (1u << 31);
    pl->lazy = p < (1u << 30);
    PrimeBarrett fb = f->prime_barrett();
    const size_t npow = n / 2 ? n / 2 : 1;
    cudaError_t e = cudaSuccess;
    if (!pl->fast) {
        e = cudaMalloc((void**)&pl->d_pow, npow * sizeof(uint32_t));
        if (e == cudaSuccess) {
            k_build_pow<<<(unsigned)((npow + 255) / 256), 256>>>(fb, pl->omega, pl->d_pow, npow);
            e = cudaGetLastError();
        }
    } else {
        // rho = omega^(N/32) is the 32nd root shared by every in-register radix stage
        const uint32_t rho = host_powmod(pl->omega, n / 32, p);
        for (int i = 0; i < 16; ++i) {
            pl->rho[i] = host_powmod(rho, i, p);
            pl->rho_sh[i] = (uint32_t)((((unsigned long long)pl->rho[i]) << 32) / p);
        }
        if (pl->log_n <= 10) {
            // single kernel, N = 32 RA: zt[k1*32 + t] = omega^(t k1), k1 < RA
            e = cudaMalloc((void**)&pl->d_zt, n * sizeof(uint2));
            if (e == cudaSuccess) {
                k_build_zt<<<(unsigned)((n + 255) / 256), 256>>>(fb, pl->omega, pl->d_zt, (int)n);
                e = cudaGetLastError();
            }
        } else {
            // N = n1 x 1024: zeta1 = omega^1024 has order n1 (column transforms, needed when n1 >= 32),
            // zeta2 = omega^n1 has order 1024 (row transforms)
            const size_t n1 = n >> 10;
            const uint32_t zeta1 = host_powmod(pl->omega, 1024, p);
            const uint32_t zeta2 = host_powmod(pl->omega, n1, p);
            e = cudaMalloc((void**)&pl->d_tw4, n * sizeof(uint2));
            if (e == cudaSuccess) e = cudaMalloc((void**)&pl->d_zt, 1024 * sizeof(uint2));
            if (e == cudaSuccess && n1 >= 32) e = cudaMalloc((void**)&pl->d_zt1, n1 * sizeof(uint2));
            if (e == cudaSuccess) {
                k_build_tw4<<<(unsigned)((n + 255) / 256), 256>>>(fb, pl->omega, pl->scale, pl->d_tw4, n);
                k_build_zt<<<4, 256>>>(fb, zeta2, pl->d_zt, 1024);
                if (n1 >= 32) k_build_zt<<<(unsigned)((n1 + 255) / 256), 256>>>(fb, zeta1, pl->d_zt1, (int)n1);
                e = cudaGetLastError();
            }
        }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_last_error("gfb_ntt_plan_create", e);
        gfb_ntt_plan_destroy(pl);
        return GFB_ERR_CUDA;
    }
    *out = pl;
    return GFB_OK;
}

void gfb_ntt_plan_destroy(gfb_ntt_plan* pl) {
    if (!pl) return;
    if (pl->d_pow) cudaFree(pl->d_pow);
    if (pl->d_tw4) cudaFree(pl->d_tw4);
    if (pl->d_zt) cudaFree(pl->d_zt);
    if (pl->d_zt1) cudaFree(pl->d_zt1);
    if (pl->d_pow_mixed) cudaFree(pl->d_pow_mixed);
    if (pl->sub) gfb_ntt_plan_destroy(pl->sub);
    free(pl);
}

size_t gfb_ntt_workspace_bytes(const gfb_ntt_plan* pl, int elem_bytes, size_t batch) {
    using namespace gfb;
    if (!pl || batch == 0) return 0;
    if (pl->mixed) {
        const size_t cap = mixed_rows_per_chunk(pl, elem_bytes);
        const size_t rows = batch < cap ? batch : cap;
        if (pl->sub)  // hybrid: two row buffers + the power-of-two plan's scratch for rows * n_odd transforms
            return 2 * rows * pl->n * (size_t)elem_bytes + gfb_ntt_workspace_bytes(pl->sub, elem_bytes, rows * pl->n_odd);
        if (pl->n_factors <= 1) return 0;  // a single stage goes straight from x to y
        return 2 * rows * pl->n * (size_t)elem_bytes;
    }
    if (pl->fast) {
        if (pl->log_n <= 10) return 0;  // single kernel
        const size_t chunk = fast_chunk(pl);
        const size_t rows = batch < chunk ? batch : chunk;
        size_t bytes = rows * pl->n * sizeof(uint32_t);
        if (pl->log_n == 20 && elem_bytes == 4) {  // the fused kernel: intermediate ring + task counters
            const size_t fb = fused::workspace_bytes(batch);
            if (fb > bytes) bytes = fb;
        }
        return bytes;
    }
    if (pl->n <= 2) return 0;
    size_t rows = kGenericWsBytes / (pl->n * sizeof(uint32_t));
    if (rows == 0) rows = 1;
    return 2 * rows * pl->n * sizeof(uint32_t);
}

gfb_status gfb_ntt_exec(gfb_ntt_plan* pl, int eb, const void* x, void* y, size_t batch, void* workspace, void* stream) {
    using namespace gfb;
    if (!pl) return GFB_ERR_BAD_ARG;
    if (batch == 0) return GFB_OK;
    if (!x || !y) return GFB_ERR_BAD_ARG;
    if (eb != 4 && eb != 8) return GFB_ERR_BAD_ARG;
    if (gfb_ntt_workspace_bytes(pl, eb, batch) != 0 && !workspace) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->mixed) {
        if (!pl->sub && pl->n_factors == 1 && x == y) return GFB_ERR_BAD_ARG;  // one out-of-place stage
        return run_mixed(pl, eb, x, y, batch, workspace, st);
    }
    if (pl->fast) {
        uint32_t* ws = (uint32_t*)workspace;
        if (eb == 4)
            return pl->lazy ? run_fast_t<uint32_t, true>(pl, (const uint32_t*)x, (uint32_t*)y, batch, ws, st)
                            : run_fast_t<uint32_t, false>(pl, (const uint32_t*)x, (uint32_t*)y, batch, ws, st);
        return pl->lazy ? run_fast_t<unsigned long long, true>(pl, (const unsigned long long*)x, (unsigned long long*)y,
                                                               batch, ws, st)
                        : run_fast_t<unsigned long long, false>(pl, (const unsigned long long*)x,
                                                                (unsigned long long*)y, batch, ws, st);
    }
    if (eb == 4) return run_generic_t<uint32_t>(pl, (const uint32_t*)x, (uint32_t*)y, batch, workspace, st);
    return run_generic_t<unsigned long long>(pl, (const unsigned long long*)x, (unsigned long long*)y, batch, workspace,
                                             st);
}

// Host pipeline: chunks of whole transforms rotate over the staging slots; each slot owns its workspace.
gfb_status gfb_ntt_exec_host(gfb_ntt_plan* pl, int eb, const void* x, void* y, size_t batch) {
    using namespace gfb;
    if (!pl) return GFB_ERR_BAD_ARG;
    if (batch == 0) return GFB_OK;
    if (!x || !y || (eb != 4 && eb != 8)) return GFB_ERR_BAD_ARG;
    HostPipeline* hp = HostPipeline::get();
    if (!hp) return GFB_ERR_CUDA;
    PipelineScope scope(hp);
    const size_t row_bytes = pl->n * (size_t)eb;
    size_t rows = kChunkBytes / row_bytes;
    if (rows == 0) rows = 1;
    if (rows > batch) rows = batch;
    const size_t ws_bytes = gfb_ntt_workspace_bytes(pl, eb, rows);
    const bool x_pinned = is_pinned_host(x), y_pinned = is_pinned_host(y);
    gfb_status status = GFB_OK;
    for (size_t c = 0, off = 0; off < batch && status == GFB_OK; ++c, off += rows) {
        const int slot = (int)(c % kSlots);
        cudaStream_t st = hp->streams[slot];
        const size_t cnt = batch - off < rows ? batch - off : rows;
        void* dx = hp->buffer(slot, 0, rows * row_bytes);
        void* dy = hp->buffer(slot, 1, rows * row_bytes);
        void* dw = ws_bytes ? hp->buffer(slot, 2, ws_bytes) : nullptr;
        if (!dx || !dy || (ws_bytes && !dw)) return GFB_ERR_CUDA;
        GFB_CUDA_TRY(hp->begin_chunk(slot));
        GFB_CUDA_TRY(hp->copy_in(slot, 0, dx, (const char*)x + off * row_bytes, cnt * row_bytes, x_pinned));
        status = gfb_ntt_exec(pl, eb, dx, dy, cnt, dw, st);
        if (status != GFB_OK) break;
        GFB_CUDA_TRY(hp->copy_out(slot, 1, (char*)y + off * row_bytes, dy, cnt * row_bytes, y_pinned));
    }
    cudaError_t e = hp->drain();
    if (e != cudaSuccess) {
        set_last_error("gfb_ntt_exec_host drain", e);
        return GFB_ERR_CUDA;
    }
    return status;
}

}  // extern "C"
