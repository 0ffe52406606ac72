// gfb_ewise_impl.cuh — the element-wise kernel templates and their launchers (overview in gfb_ewise.cu).
// Included by one translation unit per field functor (gfb_ewise_<impl>.cu); gfb_ewise.cu keeps the C-ABI entry
// points and the dispatch on the field implementation.
#pragma once
#include "gfb_common.cuh"
#include "gfb_ewise_api.cuh"

namespace gfb {

constexpr int kThreads = 256;
constexpr int kUnrollDefault = 4;
// 128-bit vectors in flight per thread and operand: 4 for the word-size fields (memory-bound, wants MLP);
// 1 for the 16-byte GF(2^m) elements, whose carry-less multiply is register-hungry and ALU-bound (occupancy matters more).
template <typename T>
struct UnrollFor {
    static constexpr int value = sizeof(T) == 16 ? 1 : kUnrollDefault;
    static constexpr int min_blocks = sizeof(T) == 16 ? 3 : 1;  // register cap for the ALU-bound 16-byte fields
};


template <typename F>
struct CheapInv {
    static constexpr bool value = false;
};
template <>
struct CheapInv<BinaryLut> {
    static constexpr bool value = true;
};

// 128-bit vectors per thread in the kernels that carry a Montgomery batch inversion (INV, DIV): one field inverse is shared by
// the elements a thread holds.  Measured on a B200 (GF(2^31-1), 1e8 int64): 4 vectors (8 elements) 2.19e11 inv/s and
// 2.01e11 div/s; 16 vectors (32 elements) 1.68e11 and 1.61e11 — the longer prefix-product chains and the lower occupancy
// cost more than the amortised inverse saves, so the depth stays at 4.
template <typename T, typename F, int VEC>
struct InvUnroll {
    static constexpr int value = kUnrollDefault;
};

template <typename F, int K>
__device__ __forceinline__ void batch_inverse(const F& f, typename F::elem (&x)[K], bool& zero_seen) {
    using E = typename F::elem;
    if constexpr (CheapInv<F>::value) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (f.is_zero(x[k])) {
                zero_seen = true;
                x[k] = f.zero();
            } else {
                x[k] = f.inv1(x[k]);
            }
        }
    } else {
        E pre[K];
        bool isz[K];
        E acc = f.one();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            isz[k] = f.is_zero(x[k]);
            zero_seen |= isz[k];
            if (isz[k]) x[k] = f.one();
            pre[k] = acc;
            acc = f.mul(acc, x[k]);
        }
        E inv = f.inv1(acc);
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {
            E r = f.mul(inv, pre[k]);
            inv = f.mul(inv, x[k]);
            x[k] = isz[k] ? f.zero() : r;
        }
    }
}

template <typename F, int OP, int MODE, int K>
__device__ __forceinline__ void apply_binary(const F& f, typename F::elem (&a)[K], typename F::elem (&b)[K],
                                             typename F::elem s, bool& zero_seen) {
    // result is left in a[]
    using E = typename F::elem;
    if constexpr (OP == GFB_DIV) {
        if constexpr (MODE == MODE_AA) {
            batch_inverse<F, K>(f, b, zero_seen);
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = f.mul(a[k], b[k]);
        } else if constexpr (MODE == MODE_AS) {
            // s already holds the inverse of the scalar divisor
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = f.mul(a[k], s);
        } else {
            batch_inverse<F, K>(f, a, zero_seen);
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = f.mul(s, a[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            E x = a[k];
            E y = (MODE == MODE_AA) ? b[k] : s;
            if constexpr (MODE == MODE_SA) {
                E t = x;
                x = y;
                y = t;
            }
            if constexpr (OP == GFB_ADD) a[k] = f.add(x, y);
            if constexpr (OP == GFB_SUB) a[k] = f.sub(x, y);
            if constexpr (OP == GFB_MUL) a[k] = f.mul(x, y);
        }
    }
}

// VEC = elements per thread-load: 16/sizeof(T) (aligned fast path) or 1 (unaligned fallback).
template <typename T, typename F, int OP, int MODE, int VEC>
__global__ void __launch_bounds__(kThreads, OP == GFB_DIV ? 1 : UnrollFor<T>::min_blocks) k_binary(F f, const T* __restrict__ a, const T* __restrict__ b,
                                                      typename F::elem s, T* __restrict__ out, size_t n, int* err) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    bool zero_seen = false;
    if constexpr (OP == GFB_DIV && MODE == MODE_AS) {
        if (f.is_zero(s)) {
            zero_seen = true;
        } else {
            s = f.inv1(s);
        }
    }
    // division carries a batch inversion: keep 4 vectors per thread there to amortise the field inverse
    constexpr int kUnroll = (OP == GFB_DIV) ? InvUnroll<T, F, VEC>::value : UnrollFor<T>::value;
    const size_t nvec = n / VEC;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int K = VEC * kUnroll;

    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
        E ea[K], eb[K];
        if constexpr (VEC * sizeof(T) == 16) {
            Pack<T> pa[kUnroll], pb[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) pa[u] = ld_stream(a + (i + u * stride) * VEC);
            if constexpr (MODE == MODE_AA) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) pb[u] = ld_stream(b + (i + u * stride) * VEC);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    ea[u * VEC + v] = to_elem<E, T>(pa[u].v[v]);
                    if constexpr (MODE == MODE_AA) eb[u * VEC + v] = to_elem<E, T>(pb[u].v[v]);
                }
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                ea[u] = to_elem<E, T>(a[i + u * stride]);
                if constexpr (MODE == MODE_AA) eb[u] = to_elem<E, T>(b[i + u * stride]);
            }
        }
        apply_binary<F, OP, MODE, K>(f, ea, eb, s, zero_seen);
        if constexpr (VEC * sizeof(T) == 16) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                Pack<T> po;
#pragma unroll
                for (int v = 0; v < VEC; ++v) po.v[v] = to_store<T, E>(ea[u * VEC + v]);
                st_stream(out + (i + u * stride) * VEC, po);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) out[i + u * stride] = to_store<T, E>(ea[u]);
        }
    }
    // remaining whole vectors, one per iteration
    for (; i < nvec; i += stride) {
        E ea[VEC], eb[VEC];
        if constexpr (VEC * sizeof(T) == 16) {
            Pack<T> pa = ld_stream(a + i * VEC), pb;
            if constexpr (MODE == MODE_AA) pb = ld_stream(b + i * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                ea[v] = to_elem<E, T>(pa.v[v]);
                if constexpr (MODE == MODE_AA) eb[v] = to_elem<E, T>(pb.v[v]);
            }
        } else {
            ea[0] = to_elem<E, T>(a[i]);
            if constexpr (MODE == MODE_AA) eb[0] = to_elem<E, T>(b[i]);
        }
        apply_binary<F, OP, MODE, VEC>(f, ea, eb, s, zero_seen);
        if constexpr (VEC * sizeof(T) == 16) {
            Pack<T> po;
#pragma unroll
            for (int v = 0; v < VEC; ++v) po.v[v] = to_store<T, E>(ea[v]);
            st_stream(out + i * VEC, po);
        } else {
            out[i] = to_store<T, E>(ea[0]);
        }
    }
    // scalar tail (n % VEC elements)
    {
        size_t t = nvec * VEC + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t < n) {
            E ea[1], eb[1];
            ea[0] = to_elem<E, T>(a[t]);
            if constexpr (MODE == MODE_AA) eb[0] = to_elem<E, T>(b[t]);
            apply_binary<F, OP, MODE, 1>(f, ea, eb, s, zero_seen);
            out[t] = to_store<T, E>(ea[0]);
        }
    }
    if (zero_seen) *err = 1;
}

template <typename F, int OP, int K>
__device__ __forceinline__ void apply_unary(const F& f, typename F::elem (&a)[K], bool& zero_seen) {
    if constexpr (OP == GFB_INV) {
        batch_inverse<F, K>(f, a, zero_seen);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if constexpr (OP == GFB_NEG) a[k] = f.neg(a[k]);
            if constexpr (OP == GFB_SQUARE) a[k] = f.mul(a[k], a[k]);
        }
    }
}

template <typename T, typename F, int OP, int VEC>
__global__ void __launch_bounds__(kThreads, OP == GFB_INV ? 1 : UnrollFor<T>::min_blocks) k_unary(F f, const T* __restrict__ a, T* __restrict__ out, size_t n,
                                                     int* err) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    bool zero_seen = false;
    // inversion amortises one field inverse over the K elements a thread holds: keep 4 vectors there
    constexpr int kUnroll = (OP == GFB_INV) ? InvUnroll<T, F, VEC>::value : UnrollFor<T>::value;
    const size_t nvec = n / VEC;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int K = VEC * kUnroll;
    for (; i + (kUnroll - 1) * stride < nvec; i += kUnroll * stride) {
        E ea[K];
        if constexpr (VEC * sizeof(T) == 16) {
            Pack<T> pa[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) pa[u] = ld_stream(a + (i + u * stride) * VEC);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) ea[u * VEC + v] = to_elem<E, T>(pa[u].v[v]);
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) ea[u] = to_elem<E, T>(a[i + u * stride]);
        }
        apply_unary<F, OP, K>(f, ea, zero_seen);
        if constexpr (VEC * sizeof(T) == 16) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                Pack<T> po;
#pragma unroll
                for (int v = 0; v < VEC; ++v) po.v[v] = to_store<T, E>(ea[u * VEC + v]);
                st_stream(out + (i + u * stride) * VEC, po);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) out[i + u * stride] = to_store<T, E>(ea[u]);
        }
    }
    for (; i < nvec; i += stride) {
        E ea[VEC];
        if constexpr (VEC * sizeof(T) == 16) {
            Pack<T> pa = ld_stream(a + i * VEC);
#pragma unroll
            for (int v = 0; v < VEC; ++v) ea[v] = to_elem<E, T>(pa.v[v]);
        } else {
            ea[0] = to_elem<E, T>(a[i]);
        }
        apply_unary<F, OP, VEC>(f, ea, zero_seen);
        if constexpr (VEC * sizeof(T) == 16) {
            Pack<T> po;
#pragma unroll
            for (int v = 0; v < VEC; ++v) po.v[v] = to_store<T, E>(ea[v]);
            st_stream(out + i * VEC, po);
        } else {
            out[i] = to_store<T, E>(ea[0]);
        }
    }
    {
        size_t t = nvec * VEC + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t < n) {
            E ea[1];
            ea[0] = to_elem<E, T>(a[t]);
            apply_unary<F, OP, 1>(f, ea, zero_seen);
            out[t] = to_store<T, E>(ea[0]);
        }
    }
    if (zero_seen) *err = 1;
}

// a ** e with per-element (or uniform) int64 exponents.
//   a == 0: e < 0 → ZeroDivisionError flag; e == 0 → 1; e > 0 → 0       (_calculate.py:581-592)
//   a != 0: a^(e mod (order-1)) when order-1 fits 64 bits; otherwise invert for e < 0.
template <typename F, bool WIDE>
__device__ __forceinline__ typename F::elem power_one(const F& f, typename F::elem a, long long e, bool& zero_seen) {
    using E = typename F::elem;
    if (e == 0) return f.one();
    if (f.is_zero(a)) {
        if (e < 0) zero_seen = true;
        return f.zero();
    }
    if constexpr (!WIDE) {
        const unsigned long long g = f.group_order();  // order - 1 >= 1
        unsigned long long r;
        if (e >= 0) {
            r = (unsigned long long)e % g;
        } else {
            unsigned long long ne = 0ull - (unsigned long long)e;  // |e|, also right for INT64_MIN
            unsigned long long t = ne % g;
            r = t ? g - t : 0ull;
        }
        return f.pow_u64(a, r);
    } else {
        if (e < 0) {
            a = f.inv1(a);
            return f.pow_u64(a, 0ull - (unsigned long long)e);
        }
        return f.pow_u64(a, (unsigned long long)e);
    }
}

template <typename T, typename F, bool SCALAR_E, bool WIDE>
__global__ void __launch_bounds__(kThreads) k_power(F f, const T* __restrict__ a, const long long* __restrict__ e,
                                                     long long e_scalar, T* __restrict__ out, size_t n, int* err) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    bool zero_seen = false;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        E x = to_elem<E, T>(a[i]);
        long long ee = SCALAR_E ? e_scalar : e[i];
        out[i] = to_store<T, E>(power_one<F, WIDE>(f, x, ee, zero_seen));
    }
    if (zero_seen) *err = 1;
}

// ----------------------------------------------------------------------------------------
// host-side dispatch
// ----------------------------------------------------------------------------------------
static inline int grid_for(size_t work_items, int blocks_per_sm) {
    size_t want = (work_items + kThreads - 1) / kThreads;
    size_t cap = (size_t)sm_count() * blocks_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename F>
static size_t smem_for(const F& f) {
    if constexpr (F::kNeedsSmem)
        return f.smem_bytes();
    else
        return 0;
}

template <typename T, typename F, int OP, int MODE>
static gfb_status launch_binary_t(F f, const void* a, const void* b, typename F::elem s, void* out, size_t n, int* err,
                                  cudaStream_t st) {
    constexpr int VEC = 16 / sizeof(T);
    const bool fast = aligned16(a) && aligned16(out) && (MODE != MODE_AA || aligned16(b));
    const size_t smem = smem_for(f);
    if (fast) {
        size_t items = (n / VEC + kUnrollDefault - 1) / kUnrollDefault + (n % VEC) + 1;
        int grid = grid_for(items, 16);
        k_binary<T, F, OP, MODE, VEC><<<grid, kThreads, smem, st>>>(f, (const T*)a, (const T*)b, s, (T*)out, n, err);
    } else {
        size_t items = (n + kUnrollDefault - 1) / kUnrollDefault + 1;
        int grid = grid_for(items, 16);
        k_binary<T, F, OP, MODE, 1><<<grid, kThreads, smem, st>>>(f, (const T*)a, (const T*)b, s, (T*)out, n, err);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename T, typename F, int MODE>
static gfb_status launch_binary_op(F f, int op, const void* a, const void* b, typename F::elem s, void* out, size_t n,
                                   int* err, cudaStream_t st) {
    switch (op) {
        case GFB_ADD: return launch_binary_t<T, F, GFB_ADD, MODE>(f, a, b, s, out, n, err, st);
        case GFB_SUB: return launch_binary_t<T, F, GFB_SUB, MODE>(f, a, b, s, out, n, err, st);
        case GFB_MUL: return launch_binary_t<T, F, GFB_MUL, MODE>(f, a, b, s, out, n, err, st);
        case GFB_DIV: return launch_binary_t<T, F, GFB_DIV, MODE>(f, a, b, s, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename T, typename F>
static gfb_status launch_binary_mode(F f, int op, int mode, const void* a, const void* b, typename F::elem s, void* out,
                                     size_t n, int* err, cudaStream_t st) {
    switch (mode) {
        case MODE_AA: return launch_binary_op<T, F, MODE_AA>(f, op, a, b, s, out, n, err, st);
        case MODE_AS: return launch_binary_op<T, F, MODE_AS>(f, op, a, b, s, out, n, err, st);
        case MODE_SA: return launch_binary_op<T, F, MODE_SA>(f, op, a, b, s, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename F>
static gfb_status launch_binary_int(F f, int op, int mode, int eb, const void* a, const void* b, uint64_t s, void* out,
                                    size_t n, int* err, cudaStream_t st) {
    using E = typename F::elem;
    switch (eb) {
        case 1: return launch_binary_mode<uint8_t, F>(f, op, mode, a, b, (E)s, out, n, err, st);
        case 2: return launch_binary_mode<uint16_t, F>(f, op, mode, a, b, (E)s, out, n, err, st);
        case 4: return launch_binary_mode<uint32_t, F>(f, op, mode, a, b, (E)s, out, n, err, st);
        case 8: return launch_binary_mode<unsigned long long, F>(f, op, mode, a, b, (E)s, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename T, typename F, int OP>
static gfb_status launch_unary_t(F f, const void* a, void* out, size_t n, int* err, cudaStream_t st) {
    constexpr int VEC = 16 / sizeof(T);
    const bool fast = aligned16(a) && aligned16(out);
    const size_t smem = smem_for(f);
    if (fast) {
        size_t items = (n / VEC + kUnrollDefault - 1) / kUnrollDefault + (n % VEC) + 1;
        k_unary<T, F, OP, VEC><<<grid_for(items, 16), kThreads, smem, st>>>(f, (const T*)a, (T*)out, n, err);
    } else {
        size_t items = (n + kUnrollDefault - 1) / kUnrollDefault + 1;
        k_unary<T, F, OP, 1><<<grid_for(items, 16), kThreads, smem, st>>>(f, (const T*)a, (T*)out, n, err);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename T, typename F>
static gfb_status launch_unary_op(F f, int op, const void* a, void* out, size_t n, int* err, cudaStream_t st) {
    switch (op) {
        case GFB_NEG: return launch_unary_t<T, F, GFB_NEG>(f, a, out, n, err, st);
        case GFB_INV: return launch_unary_t<T, F, GFB_INV>(f, a, out, n, err, st);
        case GFB_SQUARE: return launch_unary_t<T, F, GFB_SQUARE>(f, a, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename F>
static gfb_status launch_unary_int(F f, int op, int eb, const void* a, void* out, size_t n, int* err, cudaStream_t st) {
    switch (eb) {
        case 1: return launch_unary_op<uint8_t, F>(f, op, a, out, n, err, st);
        case 2: return launch_unary_op<uint16_t, F>(f, op, a, out, n, err, st);
        case 4: return launch_unary_op<uint32_t, F>(f, op, a, out, n, err, st);
        case 8: return launch_unary_op<unsigned long long, F>(f, op, a, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

template <typename T, typename F, bool WIDE>
static gfb_status launch_power_t(F f, const void* a, const int64_t* e, int64_t es, bool scalar_e, void* out, size_t n,
                                 int* err, cudaStream_t st) {
    const size_t smem = smem_for(f);
    int grid = grid_for(n, 16);
    if (scalar_e)
        k_power<T, F, true, WIDE><<<grid, kThreads, smem, st>>>(f, (const T*)a, nullptr, (long long)es, (T*)out, n, err);
    else
        k_power<T, F, false, WIDE>
            <<<grid, kThreads, smem, st>>>(f, (const T*)a, (const long long*)e, 0ll, (T*)out, n, err);
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

template <typename F>
static gfb_status launch_power_int(F f, int eb, const void* a, const int64_t* e, int64_t es, bool scalar_e, void* out,
                                   size_t n, int* err, cudaStream_t st) {
    switch (eb) {
        case 1: return launch_power_t<uint8_t, F, false>(f, a, e, es, scalar_e, out, n, err, st);
        case 2: return launch_power_t<uint16_t, F, false>(f, a, e, es, scalar_e, out, n, err, st);
        case 4: return launch_power_t<uint32_t, F, false>(f, a, e, es, scalar_e, out, n, err, st);
        case 8: return launch_power_t<unsigned long long, F, false>(f, a, e, es, scalar_e, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

// the three entry points of one functor whose elements fit 64 bits
#define GFB_EWISE_DEFINE(NAME, FUNCTOR)                                                                              \
    gfb_status ewise_binary_##NAME(FUNCTOR f, int op, int mode, int eb, const void* a, const void* b, uint64_t s_lo,  \
                                   uint64_t, void* out, size_t n, int* err, cudaStream_t st) {                       \
        return launch_binary_int(f, op, mode, eb, a, b, s_lo, out, n, err, st);                                      \
    }                                                                                                                 \
    gfb_status ewise_unary_##NAME(FUNCTOR f, int op, int eb, const void* a, void* out, size_t n, int* err,           \
                                  cudaStream_t st) {                                                                  \
        return launch_unary_int(f, op, eb, a, out, n, err, st);                                                      \
    }                                                                                                                 \
    gfb_status ewise_power_##NAME(FUNCTOR f, int eb, const void* a, const int64_t* e, int64_t es, bool scalar_e,      \
                                  void* out, size_t n, int* err, cudaStream_t st) {                                  \
        return launch_power_int(f, eb, a, e, es, scalar_e, out, n, err, st);                                         \
    }

// the three entry points of one functor at ONE storage width T
#define GFB_EWISE_DEFINE_W(NAME, FUNCTOR, T)                                                                         \
    gfb_status ewise_binary_##NAME(FUNCTOR f, int op, int mode, const void* a, const void* b, uint64_t s, void* out,  \
                                   size_t n, int* err, cudaStream_t st) {                                            \
        return launch_binary_mode<T, FUNCTOR>(f, op, mode, a, b, (typename FUNCTOR::elem)s, out, n, err, st);        \
    }                                                                                                                 \
    gfb_status ewise_unary_##NAME(FUNCTOR f, int op, const void* a, void* out, size_t n, int* err, cudaStream_t st) { \
        return launch_unary_op<T, FUNCTOR>(f, op, a, out, n, err, st);                                               \
    }                                                                                                                 \
    gfb_status ewise_power_##NAME(FUNCTOR f, const void* a, const int64_t* e, int64_t es, bool scalar_e, void* out,   \
                                  size_t n, int* err, cudaStream_t st) {                                             \
        return launch_power_t<T, FUNCTOR, false>(f, a, e, es, scalar_e, out, n, err, st);                            \
    }

}  // namespace gfb
