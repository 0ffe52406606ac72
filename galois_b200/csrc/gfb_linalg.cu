// gfb_linalg.cu — polynomial convolution and matrix multiplication over GF(p) / GF(2^m).
//
// gfb_convolve replaces convolve_jit (/root/reference/src/galois/_domains/_function.py:111-167):
//   c[k] = sum_i a[i] * b[k-i] in the field.  The reference has two code paths (integer np.convolve
//   + one `% p` when it cannot overflow, schoolbook with ADD/MULTIPLY otherwise); both compute this
//   same exact value, so one kernel serves both.
// gfb_matmul replaces matmul_jit (_domains/_linalg.py:286-308) and, for prime fields, the float-BLAS
//   detour `_lapack_linalg` (_linalg.py:21-75) — again the same exact field result.
//
// These are integer kernels; the GF(p) inner products accumulate unreduced 64-bit products in a
// 128-bit accumulator and reduce once per output.  No tensor cores (exact integer work).
#include <map>
#include <mutex>
#include <utility>

#include "gfb_common.cuh"
#include "gfb_dot.cuh"

namespace gfb {

// ----------------------------------------------------------------------------------------
// convolution: one thread per output coefficient, operands staged through shared memory tiles
// ----------------------------------------------------------------------------------------
constexpr int kConvThreads = 256;

template <typename T, typename F>
__global__ void __launch_bounds__(kConvThreads) k_convolve(F f, const T* __restrict__ a, size_t na,
                                                            const T* __restrict__ b, size_t nb, T* __restrict__ out) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    const size_t nc = na + nb - 1;
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    // i ranges over max(0, k-nb+1) .. min(k, na-1)
    const size_t i0 = k >= nb - 1 ? k - (nb - 1) : 0;
    const size_t i1 = k < na - 1 ? k : na - 1;
    DotAcc<F> acc;
    acc.init(f);
    for (size_t i = i0; i <= i1; ++i) acc.fma(f, to_elem<E, T>(a[i]), to_elem<E, T>(b[k - i]));
    out[k] = to_store<T, E>(acc.done(f));
}

// ----------------------------------------------------------------------------------------
// matmul: 16x16 output tile per block, K swept in tiles of 16 through shared memory
// ----------------------------------------------------------------------------------------
constexpr int kTile = 16;

template <typename T, typename F>
__global__ void __launch_bounds__(kTile* kTile) k_matmul(F f, const T* __restrict__ A, const T* __restrict__ B,
                                                          T* __restrict__ C, size_t M, size_t K, size_t N, int b_batched,
                                                          size_t smem_field_bytes) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    E* sA = reinterpret_cast<E*>(reinterpret_cast<char*>(smem_raw) + smem_field_bytes);
    E* sB = sA + kTile * (kTile + 1);
    const size_t batch = blockIdx.z;
    const T* Ab = A + batch * M * K;
    const T* Bb = B + (b_batched ? batch * K * N : 0);
    T* Cb = C + batch * M * N;
    const int tx = threadIdx.x % kTile, ty = threadIdx.x / kTile;
    const size_t row = (size_t)blockIdx.y * kTile + ty;
    const size_t col = (size_t)blockIdx.x * kTile + tx;
    DotAcc<F> acc;
    acc.init(f);
    for (size_t k0 = 0; k0 < K; k0 += kTile) {
        sA[ty * (kTile + 1) + tx] = (row < M && k0 + tx < K) ? to_elem<E, T>(Ab[row * K + k0 + tx]) : f.zero();
        sB[ty * (kTile + 1) + tx] = (k0 + ty < K && col < N) ? to_elem<E, T>(Bb[(k0 + ty) * N + col]) : f.zero();
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kTile; ++kk) acc.fma(f, sA[ty * (kTile + 1) + kk], sB[kk * (kTile + 1) + tx]);
        __syncthreads();
    }
    if (row < M && col < N) Cb[row * N + col] = to_store<T, E>(acc.done(f));
}

// ----------------------------------------------------------------------------------------
// polynomial evaluation at many points (Horner): y[i] = coeffs[0] x^(nc-1) + ... + coeffs[nc-1]
// evaluate_elementwise_jit.implementation, /root/reference/src/galois/_polys/_dense.py:432-440.
// One thread per point; the coefficients stream through shared memory in chunks shared by the block.
// ----------------------------------------------------------------------------------------
constexpr int kEvalThreads = 256;
constexpr int kEvalChunk = 512;

template <typename T, typename F>
__global__ void __launch_bounds__(kEvalThreads) k_poly_eval(F f, const T* __restrict__ coeffs, size_t nc,
                                                             const T* __restrict__ x, T* __restrict__ y, size_t n,
                                                             size_t smem_field_bytes) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    E* sc = reinterpret_cast<E*>(reinterpret_cast<char*>(smem_raw) + smem_field_bytes);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // every thread of the block walks the same number of rounds so that the barriers line up
    const size_t rounds = (n + stride - 1) / stride;
    for (size_t r = 0; r < rounds; ++r) {
        const size_t i = r * stride + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool live = i < n;
        const E xi = live ? to_elem<E, T>(x[i]) : f.zero();
        E acc = f.zero();
        for (size_t c0 = 0; c0 < nc; c0 += kEvalChunk) {
            const size_t cnt = nc - c0 < (size_t)kEvalChunk ? nc - c0 : (size_t)kEvalChunk;
            __syncthreads();
            for (size_t j = threadIdx.x; j < cnt; j += blockDim.x) sc[j] = to_elem<E, T>(coeffs[c0 + j]);
            __syncthreads();
            for (size_t j = 0; j < cnt; ++j) acc = f.add(sc[j], f.mul(acc, xi));
        }
        if (live) y[i] = to_store<T, E>(acc);
    }
}

template <typename F>
static size_t smem_for(const F& f) {
    if constexpr (F::kNeedsSmem)
        return (f.smem_bytes() + 15) & ~(size_t)15;
    else
        return 0;
}

template <typename T, typename F>
static gfb_status launch_convolve(F f, const void* a, size_t na, const void* b, size_t nb, void* out, cudaStream_t st) {
    const size_t nc = na + nb - 1;
    const unsigned grid = (unsigned)((nc + kConvThreads - 1) / kConvThreads);
    k_convolve<T, F><<<grid, kConvThreads, smem_for(f), st>>>(f, (const T*)a, na, (const T*)b, nb, (T*)out);
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename T, typename F>
static gfb_status launch_matmul(F f, const void* A, const void* B, void* C, size_t batch, size_t M, size_t K, size_t N,
                                int b_batched, cudaStream_t st) {
    using E = typename F::elem;
    const size_t fs = smem_for(f);
    const size_t smem = fs + 2 * kTile * (kTile + 1) * sizeof(E);
    dim3 grid((unsigned)((N + kTile - 1) / kTile), (unsigned)((M + kTile - 1) / kTile), (unsigned)batch);
    if (batch > 65535 || grid.y > 65535) return GFB_ERR_UNSUPPORTED;
    k_matmul<T, F><<<grid, kTile * kTile, smem, st>>>(f, (const T*)A, (const T*)B, (T*)C, M, K, N, b_batched, fs);
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename T, typename F>
static gfb_status launch_poly_eval(F f, const void* coeffs, size_t nc, const void* x, void* y, size_t n,
                                   cudaStream_t st) {
    using E = typename F::elem;
    const size_t fs = smem_for(f);
    const size_t smem = fs + kEvalChunk * sizeof(E);
    size_t blocks = (n + kEvalThreads - 1) / kEvalThreads;
    const size_t cap = (size_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    k_poly_eval<T, F><<<(unsigned)blocks, kEvalThreads, smem, st>>>(f, (const T*)coeffs, nc, (const T*)x, (T*)y, n, fs);
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename F>
static gfb_status eval_by_width(F f, int eb, const void* coeffs, size_t nc, const void* x, void* y, size_t n,
                                cudaStream_t st) {
    switch (eb) {
        case 1: return launch_poly_eval<uint8_t, F>(f, coeffs, nc, x, y, n, st);
        case 2: return launch_poly_eval<uint16_t, F>(f, coeffs, nc, x, y, n, st);
        case 4: return launch_poly_eval<uint32_t, F>(f, coeffs, nc, x, y, n, st);
        case 8: return launch_poly_eval<unsigned long long, F>(f, coeffs, nc, x, y, n, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename F>
static gfb_status conv_by_width(F f, int eb, const void* a, size_t na, const void* b, size_t nb, void* out,
                                cudaStream_t st) {
    switch (eb) {
        case 1: return launch_convolve<uint8_t, F>(f, a, na, b, nb, out, st);
        case 2: return launch_convolve<uint16_t, F>(f, a, na, b, nb, out, st);
        case 4: return launch_convolve<uint32_t, F>(f, a, na, b, nb, out, st);
        case 8: return launch_convolve<unsigned long long, F>(f, a, na, b, nb, out, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename F>
static gfb_status matmul_by_width(F f, int eb, const void* A, const void* B, void* C, size_t batch, size_t M, size_t K,
                                  size_t N, int bb, cudaStream_t st) {
    switch (eb) {
        case 1: return launch_matmul<uint8_t, F>(f, A, B, C, batch, M, K, N, bb, st);
        case 2: return launch_matmul<uint16_t, F>(f, A, B, C, batch, M, K, N, bb, st);
        case 4: return launch_matmul<uint32_t, F>(f, A, B, C, batch, M, K, N, bb, st);
        case 8: return launch_matmul<unsigned long long, F>(f, A, B, C, batch, M, K, N, bb, st);
    }
    return GFB_ERR_BAD_ARG;
}

// ----------------------------------------------------------------------------------------
// NTT-based convolution over GF(p) when the padded length L = 2^k divides p - 1 (e.g. p = 998244353):
//   c = INTT_L( NTT_L(a padded) .* NTT_L(b padded) ),  L >= na + nb - 1.
// The field result is unique, so this is bit-identical to the reference's schoolbook / np.convolve paths
// (_domains/_function.py:136-167) whichever primitive L-th root is used; it replaces O(na nb) work by three
// transforms on the kernels of gfb_ntt.cu.
// ----------------------------------------------------------------------------------------
template <typename T>
__global__ void k_conv_pad(const T* __restrict__ a, size_t na, const T* __restrict__ b, size_t nb, uint32_t* __restrict__ AB,
                           size_t L) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * L) return;
    const size_t j = i < L ? i : i - L;
    uint32_t v = 0;
    if (i < L) {
        if (j < na) v = (uint32_t)a[j];
    } else if (j < nb) {
        v = (uint32_t)b[j];
    }
    AB[i] = v;
}

template <typename T>
__global__ void k_conv_unpad(const uint32_t* __restrict__ A, T* __restrict__ out, size_t nc) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nc) out[i] = (T)A[i];
}

struct ConvPlans {
    gfb_ntt_plan* fwd;
    gfb_ntt_plan* inv;
};
static std::mutex g_conv_mu;
static std::map<std::pair<gfb_field*, size_t>, ConvPlans> g_conv_plans;

static uint32_t host_pow_u32(uint32_t b, unsigned long long e, uint32_t p) {
    unsigned long long r = 1 % p, x = b % p;
    while (e) {
        if (e & 1) r = r * x % p;
        x = x * x % p;
        e >>= 1;
    }
    return (uint32_t)r;
}

void conv_cache_drop(gfb_field* f) {
    std::lock_guard<std::mutex> lk(g_conv_mu);
    for (auto it = g_conv_plans.begin(); it != g_conv_plans.end();) {
        if (it->first.first == f) {
            gfb_ntt_plan_destroy(it->second.fwd);
            gfb_ntt_plan_destroy(it->second.inv);
            it = g_conv_plans.erase(it);
        } else {
            ++it;
        }
    }
}

// padded transform length, or 0 when the NTT route does not apply
static size_t conv_ntt_length(const gfb_field* f, size_t na, size_t nb) {
    if (f->kind != GFB_PRIME || f->p < 3 || f->p >= (1u << 31)) return 0;
    const size_t nc = na + nb - 1;
    const size_t small = na < nb ? na : nb;
    if (small < 32 || nc < 1024) return 0;  // the schoolbook kernel wins on short operands
    size_t L = 1;
    while (L < nc) L <<= 1;
    if (L > ((size_t)1 << 24) || ((unsigned long long)f->p - 1ull) % L != 0) return 0;
    return L;
}

template <typename T>
static gfb_status convolve_ntt_t(gfb_field* f, const T* a, size_t na, const T* b, size_t nb, T* out, size_t L,
                                 cudaStream_t st) {
    ConvPlans plans;
    {
        std::lock_guard<std::mutex> lk(g_conv_mu);
        auto key = std::make_pair(f, L);
        auto it = g_conv_plans.find(key);
        if (it == g_conv_plans.end()) {
            const uint32_t p = f->p;
            const uint32_t omega = host_pow_u32((uint32_t)f->primitive_element, ((unsigned long long)p - 1ull) / L, p);
            const uint32_t omega_inv = host_pow_u32(omega, (unsigned long long)p - 2ull, p);
            const uint32_t scale = host_pow_u32((uint32_t)(L % p), (unsigned long long)p - 2ull, p);
            ConvPlans np{nullptr, nullptr};
            gfb_status rc = gfb_ntt_plan_create(&np.fwd, f, L, omega, 1);
            if (rc == GFB_OK) rc = gfb_ntt_plan_create(&np.inv, f, L, omega_inv, scale);
            if (rc != GFB_OK) {
                gfb_ntt_plan_destroy(np.fwd);
                gfb_ntt_plan_destroy(np.inv);
                return rc;
            }
            it = g_conv_plans.emplace(key, np).first;
        }
        plans = it->second;
    }
    const size_t nc = na + nb - 1;
    size_t ws_bytes = gfb_ntt_workspace_bytes(plans.fwd, 4, 2);
    const size_t ws_inv = gfb_ntt_workspace_bytes(plans.inv, 4, 1);
    if (ws_inv > ws_bytes) ws_bytes = ws_inv;
    char* mem = nullptr;
    GFB_CUDA_TRY(cudaMallocAsync((void**)&mem, 2 * L * sizeof(uint32_t) + ws_bytes + 256, st));
    uint32_t* AB = (uint32_t*)mem;
    void* ws = ws_bytes ? (void*)(mem + ((2 * L * sizeof(uint32_t) + 255) & ~(size_t)255)) : nullptr;
    gfb_status rc = GFB_OK;
    k_conv_pad<T><<<(unsigned)((2 * L + 255) / 256), 256, 0, st>>>(a, na, b, nb, AB, L);
    rc = gfb_ntt_exec(plans.fwd, 4, AB, AB, 2, ws, st);
    if (rc == GFB_OK) rc = gfb_binary(f, GFB_MUL, 4, AB, AB + L, AB, L, st);
    if (rc == GFB_OK) rc = gfb_ntt_exec(plans.inv, 4, AB, AB, 1, ws, st);
    if (rc == GFB_OK) {
        k_conv_unpad<T><<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(AB, out, nc);
        if (cudaGetLastError() != cudaSuccess) rc = GFB_ERR_CUDA;
    }
    cudaFreeAsync(mem, st);
    return rc;
}

static gfb_status convolve_ntt(gfb_field* f, int eb, const void* a, size_t na, const void* b, size_t nb, void* out,
                               size_t L, cudaStream_t st) {
    switch (eb) {
        case 1: return convolve_ntt_t<uint8_t>(f, (const uint8_t*)a, na, (const uint8_t*)b, nb, (uint8_t*)out, L, st);
        case 2: return convolve_ntt_t<uint16_t>(f, (const uint16_t*)a, na, (const uint16_t*)b, nb, (uint16_t*)out, L, st);
        case 4: return convolve_ntt_t<uint32_t>(f, (const uint32_t*)a, na, (const uint32_t*)b, nb, (uint32_t*)out, L, st);
        case 8:
            return convolve_ntt_t<unsigned long long>(f, (const unsigned long long*)a, na, (const unsigned long long*)b, nb,
                                                      (unsigned long long*)out, L, st);
    }
    return GFB_ERR_BAD_ARG;
}

static bool width_ok(const gfb_field* f, int eb) {
    if (f->impl == IMPL_BIN_128) return eb == 16;
    if (eb != 1 && eb != 2 && eb != 4 && eb != 8) return false;
    if (eb == 8) return true;
    if (f->kind == GFB_PRIME) return (f->order_lo - 1) < (1ull << (8 * eb));
    return f->degree <= 8 * eb;
}

}  // namespace gfb

extern "C" {

gfb_status gfb_convolve(gfb_field* f, int eb, const void* a, size_t na, const void* b, size_t nb, void* out,
                        void* stream) {
    using namespace gfb;
    if (!f || !a || !b || !out || na == 0 || nb == 0) return GFB_ERR_BAD_ARG;
    if (!width_ok(f, eb)) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (const size_t L = conv_ntt_length(f, na, nb)) return convolve_ntt(f, eb, a, na, b, nb, out, L, st);
    switch (f->impl) {
        case IMPL_PRIME_M31: return conv_by_width(f->prime_m31(), eb, a, na, b, nb, out, st);
        case IMPL_PRIME_BARRETT: return conv_by_width(f->prime_barrett(), eb, a, na, b, nb, out, st);
        case IMPL_BIN_LUT: return conv_by_width(f->bin_lut(), eb, a, na, b, nb, out, st);
        case IMPL_BIN_CALC: return conv_by_width(f->bin_calc(), eb, a, na, b, nb, out, st);
        case IMPL_BIN_128: return launch_convolve<u128, Binary128>(f->bin_128(), a, na, b, nb, out, st);
    }
    return GFB_ERR_BAD_ARG;
}

gfb_status gfb_matmul(gfb_field* f, int eb, const void* A, const void* B, void* C, size_t batch, size_t M, size_t K,
                      size_t N, int b_batched, void* stream) {
    using namespace gfb;
    if (!f || !A || !B || !C) return GFB_ERR_BAD_ARG;
    if (batch == 0 || M == 0 || N == 0) return GFB_OK;
    if (!width_ok(f, eb)) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (f->impl) {
        case IMPL_PRIME_M31: return matmul_by_width(f->prime_m31(), eb, A, B, C, batch, M, K, N, b_batched, st);
        case IMPL_PRIME_BARRETT: return matmul_by_width(f->prime_barrett(), eb, A, B, C, batch, M, K, N, b_batched, st);
        case IMPL_BIN_LUT: return matmul_by_width(f->bin_lut(), eb, A, B, C, batch, M, K, N, b_batched, st);
        case IMPL_BIN_CALC: return matmul_by_width(f->bin_calc(), eb, A, B, C, batch, M, K, N, b_batched, st);
        case IMPL_BIN_128: return launch_matmul<u128, Binary128>(f->bin_128(), A, B, C, batch, M, K, N, b_batched, st);
    }
    return GFB_ERR_BAD_ARG;
}

gfb_status gfb_poly_evaluate(gfb_field* f, int eb, const void* coeffs, size_t nc, const void* x, void* y, size_t n,
                             void* stream) {
    using namespace gfb;
    if (!f || !coeffs || nc == 0) return GFB_ERR_BAD_ARG;
    if (n == 0) return GFB_OK;
    if (!x || !y) return GFB_ERR_BAD_ARG;
    if (!width_ok(f, eb)) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (f->impl) {
        case IMPL_PRIME_M31: return eval_by_width(f->prime_m31(), eb, coeffs, nc, x, y, n, st);
        case IMPL_PRIME_BARRETT: return eval_by_width(f->prime_barrett(), eb, coeffs, nc, x, y, n, st);
        case IMPL_BIN_LUT: return eval_by_width(f->bin_lut(), eb, coeffs, nc, x, y, n, st);
        case IMPL_BIN_CALC: return eval_by_width(f->bin_calc(), eb, coeffs, nc, x, y, n, st);
        case IMPL_BIN_128: return launch_poly_eval<u128, Binary128>(f->bin_128(), coeffs, nc, x, y, n, st);
    }
    return GFB_ERR_BAD_ARG;
}

gfb_status gfb_poly_evaluate_host(gfb_field* f, int eb, const void* coeffs, size_t nc, const void* x, void* y,
                                  size_t n) {
    if (!f || !coeffs || nc == 0) return GFB_ERR_BAD_ARG;
    if (n == 0) return GFB_OK;
    if (!x || !y) return GFB_ERR_BAD_ARG;
    char *dc = nullptr, *dx = nullptr, *dy = nullptr;
    GFB_CUDA_TRY(cudaMalloc((void**)&dc, nc * (size_t)eb + 16));
    cudaError_t e = cudaMalloc((void**)&dx, n * (size_t)eb + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dy, n * (size_t)eb + 16);
    gfb_status st = GFB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(dc, coeffs, nc * (size_t)eb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, n * (size_t)eb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = gfb_poly_evaluate(f, eb, dc, nc, dx, dy, n, nullptr);
        if (st == GFB_OK) e = cudaMemcpy(y, dy, n * (size_t)eb, cudaMemcpyDeviceToHost);
    }
    cudaFree(dc);
    cudaFree(dx);
    cudaFree(dy);
    if (e != cudaSuccess) {
        gfb::set_last_error("gfb_poly_evaluate_host", e);
        return GFB_ERR_CUDA;
    }
    return st;
}

// host variants: one-shot copies (these operands are small next to the element-wise arrays)
gfb_status gfb_convolve_host(gfb_field* f, int eb, const void* a, size_t na, const void* b, size_t nb, void* out) {
    if (!f || !a || !b || !out || na == 0 || nb == 0) return GFB_ERR_BAD_ARG;
    const size_t nc = na + nb - 1;
    char* d = nullptr;
    GFB_CUDA_TRY(cudaMalloc((void**)&d, (na + nb + nc) * (size_t)eb + 64));
    char* da = d;
    char* db = da + ((na * eb + 15) & ~(size_t)15);
    char* dc = db + ((nb * eb + 15) & ~(size_t)15);
    gfb_status st = GFB_OK;
    cudaError_t e = cudaMemcpy(da, a, na * eb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, b, nb * eb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = gfb_convolve(f, eb, da, na, db, nb, dc, nullptr);
        if (st == GFB_OK) e = cudaMemcpy(out, dc, nc * eb, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (e != cudaSuccess) {
        gfb::set_last_error("gfb_convolve_host", e);
        return GFB_ERR_CUDA;
    }
    return st;
}

gfb_status gfb_matmul_host(gfb_field* f, int eb, const void* A, const void* B, void* C, size_t batch, size_t M, size_t K,
                           size_t N, int b_batched) {
    if (!f || !A || !B || !C) return GFB_ERR_BAD_ARG;
    if (batch == 0 || M == 0 || N == 0) return GFB_OK;
    const size_t ba = batch * M * K * eb, bb = (b_batched ? batch : 1) * K * N * eb, bc = batch * M * N * eb;
    char *da = nullptr, *db = nullptr, *dc = nullptr;
    GFB_CUDA_TRY(cudaMalloc((void**)&da, ba + 16));
    cudaError_t e = cudaMalloc((void**)&db, bb + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dc, bc + 16);
    gfb_status st = GFB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(da, A, ba, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(db, B, bb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = gfb_matmul(f, eb, da, db, dc, batch, M, K, N, b_batched, nullptr);
        if (st == GFB_OK) e = cudaMemcpy(C, dc, bc, cudaMemcpyDeviceToHost);
    }
    cudaFree(da);
    cudaFree(db);
    cudaFree(dc);
    if (e != cudaSuccess) {
        gfb::set_last_error("gfb_matmul_host", e);
        return GFB_ERR_CUDA;
    }
    return st;
}

}  // extern "C"
