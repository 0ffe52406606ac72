// gfb_fold.cu — the ufunc METHODS reduce / accumulate / reduceat for field arrays on the GPU.
//
// The reference forwards every ufunc method to the compiled Numba DUFunc (`getattr(self.ufunc, method)`,
// /root/reference/src/galois/_domains/_ufunc.py:349,403) and tests them (tests/fields/test_numpy_ufuncs.py): a serial fold
// with the field's add / multiply per output.  Field add and multiply are associative and commutative, so any grouping
// gives the reference's value; the subtract / divide folds (a0 - a1 - a2 ... = a0 - (a1 + a2 + ...)) are expressed
// through them by the caller.
//
//   gfb_fold           a viewed as [outer][len][inner] (C order, any axis of an N-D array):
//                        reduce      out[outer][inner]      = fold over len
//                        accumulate  out[outer][len][inner] = running fold over len
//                      many lines  -> one thread per line, lanes along `inner` (coalesced);
//                      few long contiguous lines (inner == 1) -> chunks: per-chunk partials (block reduce), a scan of the
//                      partials, and for accumulate a second pass that re-folds each chunk from its carry-in.
//   gfb_fold_segments  np.<ufunc>.reduceat: out[s] = fold of a[start[s] .. end[s]) (one thread per segment).
#include "gfb_common.cuh"
#include "gfb_dispatch.cuh"

namespace gfb {

constexpr int kFoldThreads = 256;
constexpr int kFoldItems = 8;                                   // consecutive elements per thread in the chunked kernels
constexpr size_t kFoldChunk = (size_t)kFoldThreads * kFoldItems;  // elements per chunk

__device__ __forceinline__ uint32_t shfl_down_elem(uint32_t v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
__device__ __forceinline__ unsigned long long shfl_down_elem(unsigned long long v, int o) {
    return __shfl_down_sync(0xffffffffu, v, o);
}
__device__ __forceinline__ uint32_t shfl_up_elem(uint32_t v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
__device__ __forceinline__ unsigned long long shfl_up_elem(unsigned long long v, int o) {
    return __shfl_up_sync(0xffffffffu, v, o);
}

template <typename F, int OP>
__device__ __forceinline__ typename F::elem fold_op(const F& f, typename F::elem x, typename F::elem y) {
    if constexpr (OP == GFB_ADD)
        return f.add(x, y);
    else
        return f.mul(x, y);
}
template <typename F, int OP>
__device__ __forceinline__ typename F::elem fold_identity(const F& f) {
    if constexpr (OP == GFB_ADD)
        return f.zero();
    else
        return f.one();
}

// one thread per line (outer x inner lines), sequential over len
template <typename T, typename F, int OP, bool SCAN>
__global__ void __launch_bounds__(kFoldThreads)
    k_fold_lines(F f, const T* __restrict__ a, T* __restrict__ out, size_t outer, size_t len, size_t inner, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    const size_t lines = outer * inner;
    for (size_t line = (size_t)blockIdx.x * blockDim.x + threadIdx.x; line < lines; line += (size_t)gridDim.x * blockDim.x) {
        const size_t o = line / inner, i = line - o * inner;
        const T* src = a + o * len * inner + i;
        T* dst = SCAN ? out + o * len * inner + i : out + line;
        E acc = to_elem<E, T>(src[0]);
        if (SCAN) dst[0] = to_store<T, E>(acc);
        for (size_t j = 1; j < len; ++j) {
            acc = fold_op<F, OP>(f, acc, to_elem<E, T>(src[j * inner]));
            if (SCAN) dst[j * inner] = to_store<T, E>(acc);
        }
        if (!SCAN) dst[0] = to_store<T, E>(acc);
    }
}

// block-wide fold of one value per thread; the result is valid in thread 0 (and in every thread of warp 0)
template <typename F, int OP>
__device__ __forceinline__ typename F::elem block_fold(const F& f, typename F::elem v, typename F::elem* warp_tot) {
    using E = typename F::elem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const E other = shfl_down_elem(v, o);
        v = fold_op<F, OP>(f, v, other);
    }
    if (lane == 0) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < kFoldThreads / 32 ? warp_tot[lane] : fold_identity<F, OP>(f);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const E other = shfl_down_elem(v, o);
            v = fold_op<F, OP>(f, v, other);
        }
    }
    return v;
}

// contiguous rows: partial[row][chunk] = fold of the chunk
template <typename T, typename F, int OP>
__global__ void __launch_bounds__(kFoldThreads)
    k_fold_row_partial(F f, const T* __restrict__ a, T* __restrict__ partial, size_t len, size_t n_chunks, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    E* warp_tot = reinterpret_cast<E*>(reinterpret_cast<char*>(smem_raw) + fs);
    const size_t row = blockIdx.y, chunk = blockIdx.x;
    const T* src = a + row * len;
    const size_t lo = chunk * kFoldChunk;
    const size_t hi = lo + kFoldChunk < len ? lo + kFoldChunk : len;
    E acc = fold_identity<F, OP>(f);
    for (size_t j = lo + threadIdx.x; j < hi; j += kFoldThreads) acc = fold_op<F, OP>(f, acc, to_elem<E, T>(src[j]));
    acc = block_fold<F, OP>(f, acc, warp_tot);
    if (threadIdx.x == 0) partial[row * n_chunks + chunk] = to_store<T, E>(acc);
}

// contiguous rows, accumulate: re-fold chunk `c` from the inclusive scan of the partials (carry-in = incl[c - 1])
template <typename T, typename F, int OP>
__global__ void __launch_bounds__(kFoldThreads)
    k_fold_row_apply(F f, const T* __restrict__ a, const T* __restrict__ incl, T* __restrict__ out, size_t len, size_t n_chunks,
                     size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    E* warp_tot = reinterpret_cast<E*>(reinterpret_cast<char*>(smem_raw) + fs);
    const size_t row = blockIdx.y, chunk = blockIdx.x;
    const T* src = a + row * len;
    T* dst = out + row * len;
    const size_t base = chunk * kFoldChunk + (size_t)threadIdx.x * kFoldItems;
    E v[kFoldItems];
    E tot = fold_identity<F, OP>(f);
#pragma unroll
    for (int k = 0; k < kFoldItems; ++k) {
        v[k] = base + k < len ? to_elem<E, T>(src[base + k]) : fold_identity<F, OP>(f);
        tot = fold_op<F, OP>(f, tot, v[k]);
        v[k] = tot;  // inclusive fold inside the thread
    }
    // exclusive scan of the thread totals across the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    E incl_w = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const E other = shfl_up_elem(incl_w, o);
        if (lane >= o) incl_w = fold_op<F, OP>(f, other, incl_w);
    }
    if (lane == 31) warp_tot[warp] = incl_w;
    __syncthreads();
    E carry = chunk > 0 ? to_elem<E, T>(incl[row * n_chunks + chunk - 1]) : fold_identity<F, OP>(f);
    for (int wq = 0; wq < warp; ++wq) carry = fold_op<F, OP>(f, carry, warp_tot[wq]);
    E excl = shfl_up_elem(incl_w, 1);
    if (lane == 0) excl = fold_identity<F, OP>(f);
    carry = fold_op<F, OP>(f, carry, excl);
#pragma unroll
    for (int k = 0; k < kFoldItems; ++k)
        if (base + k < len) dst[base + k] = to_store<T, E>(fold_op<F, OP>(f, carry, v[k]));
}

// reduceat: out[s] = fold of a[start[s] .. end[s])  (end > start)
template <typename T, typename F, int OP>
__global__ void __launch_bounds__(kFoldThreads)
    k_fold_segments(F f, const T* __restrict__ a, const long long* __restrict__ start, const long long* __restrict__ end,
                    T* __restrict__ out, size_t n_seg, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += (size_t)gridDim.x * blockDim.x) {
        const long long lo = start[s], hi = end[s];
        E acc = to_elem<E, T>(a[lo]);
        for (long long j = lo + 1; j < hi; ++j) acc = fold_op<F, OP>(f, acc, to_elem<E, T>(a[j]));
        out[s] = to_store<T, E>(acc);
    }
}

static unsigned fold_grid(size_t items) {
    size_t blocks = (items + kFoldThreads - 1) / kFoldThreads;
    const size_t cap = (size_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

template <typename T, typename F, int OP>
static gfb_status fold_typed(F f, int scan, const T* a, T* out, size_t outer, size_t len, size_t inner, void* workspace,
                             cudaStream_t st) {
    using E = typename F::elem;
    const size_t fs = field_smem_bytes(f);
    const size_t lines = outer * inner;
    const bool chunked = inner == 1 && len >= 4 * kFoldChunk && lines < 2048;
    if (!chunked) {
        if (scan)
            k_fold_lines<T, F, OP, true><<<fold_grid(lines), kFoldThreads, fs, st>>>(f, a, out, outer, len, inner, fs);
        else
            k_fold_lines<T, F, OP, false><<<fold_grid(lines), kFoldThreads, fs, st>>>(f, a, out, outer, len, inner, fs);
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    }
    const size_t n_chunks = (len + kFoldChunk - 1) / kFoldChunk;
    if (n_chunks > 0x7fffffffull || outer > 65535) return GFB_ERR_UNSUPPORTED;
    T* partial = (T*)workspace;              // [outer][n_chunks]
    T* incl = partial + outer * n_chunks;    // inclusive scan of the partials
    const size_t smem = fs + (kFoldThreads / 32) * sizeof(E);
    dim3 grid((unsigned)n_chunks, (unsigned)outer);
    k_fold_row_partial<T, F, OP><<<grid, kFoldThreads, smem, st>>>(f, a, partial, len, n_chunks, fs);
    if (!scan) {
        k_fold_lines<T, F, OP, false><<<fold_grid(outer), kFoldThreads, fs, st>>>(f, partial, out, outer, n_chunks, 1, fs);
    } else {
        k_fold_lines<T, F, OP, true><<<fold_grid(outer), kFoldThreads, fs, st>>>(f, partial, incl, outer, n_chunks, 1, fs);
        k_fold_row_apply<T, F, OP><<<grid, kFoldThreads, smem, st>>>(f, a, incl, out, len, n_chunks, fs);
    }
    GFB_CUDA_TRY(cudaGetLastError());
    return GFB_OK;
}

}  // namespace gfb

extern "C" {

size_t gfb_fold_workspace_bytes(int elem_bytes, size_t outer, size_t len, size_t inner) {
    using namespace gfb;
    if (inner != 1 || len < 4 * kFoldChunk || outer * inner >= 2048) return 0;
    const size_t n_chunks = (len + kFoldChunk - 1) / kFoldChunk;
    return 2 * outer * n_chunks * (size_t)elem_bytes + 64;
}

gfb_status gfb_fold(gfb_field* f, int op, int scan, int eb, const void* a, void* out, size_t outer, size_t len, size_t inner,
                    void* workspace, void* stream) {
    using namespace gfb;
    if (!f || (op != GFB_ADD && op != GFB_MUL)) return GFB_ERR_BAD_ARG;
    if (outer == 0 || inner == 0) return GFB_OK;
    if (len == 0 || !a || !out) return GFB_ERR_BAD_ARG;
    if (gfb_fold_workspace_bytes(eb, outer, len, inner) != 0 && !workspace) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_field(f, eb, [&](auto functor, auto tag) -> gfb_status {
        using T = typename decltype(tag)::type;
        using F = decltype(functor);
        if (op == GFB_ADD) return fold_typed<T, F, GFB_ADD>(functor, scan, (const T*)a, (T*)out, outer, len, inner, workspace, st);
        return fold_typed<T, F, GFB_MUL>(functor, scan, (const T*)a, (T*)out, outer, len, inner, workspace, st);
    });
}

gfb_status gfb_fold_segments(gfb_field* f, int op, int eb, const void* a, const int64_t* start, const int64_t* end, void* out,
                             size_t n_seg, void* stream) {
    using namespace gfb;
    if (!f || (op != GFB_ADD && op != GFB_MUL)) return GFB_ERR_BAD_ARG;
    if (n_seg == 0) return GFB_OK;
    if (!a || !start || !end || !out) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_field(f, eb, [&](auto functor, auto tag) -> gfb_status {
        using T = typename decltype(tag)::type;
        using F = decltype(functor);
        const size_t fs = field_smem_bytes(functor);
        if (op == GFB_ADD)
            k_fold_segments<T, F, GFB_ADD><<<fold_grid(n_seg), kFoldThreads, fs, st>>>(
                functor, (const T*)a, (const long long*)start, (const long long*)end, (T*)out, n_seg, fs);
        else
            k_fold_segments<T, F, GFB_MUL><<<fold_grid(n_seg), kFoldThreads, fs, st>>>(
                functor, (const T*)a, (const long long*)start, (const long long*)end, (T*)out, n_seg, fs);
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    });
}

// host variants: one-shot copies around the device entry points (the operands of a fold are read once)
gfb_status gfb_fold_host(gfb_field* f, int op, int scan, int eb, const void* a, void* out, size_t outer, size_t len,
                         size_t inner) {
    using namespace gfb;
    if (!f) return GFB_ERR_BAD_ARG;
    if (outer == 0 || inner == 0) return GFB_OK;
    if (len == 0 || !a || !out) return GFB_ERR_BAD_ARG;
    const size_t in_bytes = outer * len * inner * (size_t)eb;
    const size_t out_bytes = (scan ? outer * len * inner : outer * inner) * (size_t)eb;
    const size_t ws_bytes = gfb_fold_workspace_bytes(eb, outer, len, inner);
    char *da = nullptr, *dout = nullptr, *dws = nullptr;
    cudaError_t e = cudaMalloc((void**)&da, in_bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, out_bytes + 16);
    if (e == cudaSuccess && ws_bytes) e = cudaMalloc((void**)&dws, ws_bytes);
    gfb_status rc = GFB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(da, a, in_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = gfb_fold(f, op, scan, eb, da, dout, outer, len, inner, dws, nullptr);
        if (rc == GFB_OK) e = cudaMemcpy(out, dout, out_bytes, cudaMemcpyDeviceToHost);
    }
    cudaFree(da);
    cudaFree(dout);
    cudaFree(dws);
    if (e != cudaSuccess) {
        set_last_error("gfb_fold_host", e);
        return GFB_ERR_CUDA;
    }
    return rc;
}

gfb_status gfb_fold_segments_host(gfb_field* f, int op, int eb, const void* a, size_t n, const int64_t* start,
                                  const int64_t* end, void* out, size_t n_seg) {
    using namespace gfb;
    if (!f) return GFB_ERR_BAD_ARG;
    if (n_seg == 0) return GFB_OK;
    if (!a || !start || !end || !out || n == 0) return GFB_ERR_BAD_ARG;
    for (size_t s = 0; s < n_seg; ++s)
        if (start[s] < 0 || end[s] <= start[s] || (size_t)end[s] > n) return GFB_ERR_BAD_ARG;
    char *da = nullptr, *dout = nullptr;
    int64_t* didx = nullptr;
    cudaError_t e = cudaMalloc((void**)&da, n * (size_t)eb + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, n_seg * (size_t)eb + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&didx, 2 * n_seg * sizeof(int64_t));
    gfb_status rc = GFB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(da, a, n * (size_t)eb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(didx, start, n_seg * sizeof(int64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(didx + n_seg, end, n_seg * sizeof(int64_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rc = gfb_fold_segments(f, op, eb, da, didx, didx + n_seg, dout, n_seg, nullptr);
        if (rc == GFB_OK) e = cudaMemcpy(out, dout, n_seg * (size_t)eb, cudaMemcpyDeviceToHost);
    }
    cudaFree(da);
    cudaFree(dout);
    cudaFree(didx);
    if (e != cudaSuccess) {
        set_last_error("gfb_fold_segments_host", e);
        return GFB_ERR_CUDA;
    }
    return rc;
}

}  // extern "C"
