// gfb_poly.cu — roots of a polynomial over GF(p) / GF(2^m) by exhaustive evaluation.
//
// gfb_poly_roots replaces roots_jit (/root/reference/src/galois/_polys/_dense.py:443-513).  The reference walks
// x = 0, 1, alpha, alpha^2, ... serially (Chien search: lambda_j *= alpha^degree_j per step), collects the x with
// f(x) = 0, and returns them sorted (`roots[np.argsort(roots)]`, :464-466).  The set of roots does not depend on
// the visiting order, so the device evaluates f at every element IN INTEGER ORDER, one thread per element, and
// compacts the zeros in that order — which is already the reference's sorted output:
//
//   pass 1  k_roots<false>   evaluate, count the zeros of each 4096-element tile          (order * terms multiplies)
//   pass 2  k_roots_scan     exclusive scan of the tile counts (one CTA)
//   pass 3  k_roots<true>    tiles with a non-zero count re-evaluate and write their zeros at the scanned offset
//
// The polynomial is sparse: f(x) = sum_j c_j x^(d_j), d_0 > d_1 > ...; Horner over the non-zero terms multiplies by
// x^(d_{j-1} - d_j) (a plain multiply when the gap is 1, i.e. the dense case).
#include <type_traits>

#include "gfb_common.cuh"
#include "gfb_dispatch.cuh"

namespace gfb {

constexpr int kRootsThreads = 256;
constexpr int kRootsRounds = 16;                             // elements per thread
constexpr int kRootsTile = kRootsThreads * kRootsRounds;     // elements per CTA
constexpr int kRootsChunk = 512;                             // terms staged in shared memory at a time

template <typename T, typename F, bool WRITE, bool DENSE>
__global__ void __launch_bounds__(kRootsThreads)
k_roots(F f, const T* __restrict__ coeffs, const unsigned long long* __restrict__ gaps, size_t n_terms,
        unsigned long long tail_exp, int tail_positive, unsigned long long order, unsigned long long tile0,
        unsigned int* __restrict__ counts, const unsigned long long* __restrict__ offsets, T* __restrict__ roots,
        unsigned long long max_roots, size_t smem_field_bytes) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    char* base = reinterpret_cast<char*>(smem_raw) + smem_field_bytes;
    unsigned long long* sg = reinterpret_cast<unsigned long long*>(base);
    E* sc = reinterpret_cast<E*>(base + kRootsChunk * sizeof(unsigned long long));
    __shared__ unsigned int s_warp[kRootsThreads / 32];
    __shared__ unsigned long long s_pos;

    const unsigned long long tile = tile0 + blockIdx.x;
    if (WRITE) {
        if (counts[tile] == 0u) return;  // uniform per block
        if (threadIdx.x == 0) s_pos = offsets[tile];
    }
    const bool one_chunk = n_terms <= (size_t)kRootsChunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int my_count = 0;

    for (int r = 0; r < kRootsRounds; ++r) {
        const unsigned long long xi = tile * kRootsTile + (unsigned long long)r * kRootsThreads + threadIdx.x;
        const bool live = xi < order;
        const E x = live ? (E)xi : f.one();
        E acc = f.zero();
        for (size_t c0 = 0; c0 < n_terms; c0 += kRootsChunk) {
            const size_t cnt = n_terms - c0 < (size_t)kRootsChunk ? n_terms - c0 : (size_t)kRootsChunk;
            if (!one_chunk || r == 0) {
                __syncthreads();
                for (size_t j = threadIdx.x; j < cnt; j += blockDim.x) {
                    sc[j] = to_elem<E, T>(coeffs[c0 + j]);
                    sg[j] = gaps[c0 + j];
                }
                __syncthreads();
            }
            if constexpr (DENSE) {
                // every gap is 1 (all coefficients present): plain Horner, acc = acc * x + c_j, x in multiplier form
                const E xp = f.premul(x);
#pragma unroll 8
                for (size_t j = 0; j < cnt; ++j) acc = f.add(f.mul_prepared(acc, xp), sc[j]);
            } else {
                // acc = acc * x^gap_j + c_j   (gap_0 = 0 contributes acc = c_0)
                for (size_t j = 0; j < cnt; ++j) {
                    const unsigned long long g = sg[j];
                    const E xg = g == 1ull ? x : f.pow_u64(x, g);
                    acc = f.add(f.mul(acc, xg), sc[j]);
                }
            }
        }
        bool is_root;
        if (xi == 0ull) {
            // f(0) is the constant term: zero exactly when the lowest non-zero degree is positive (:493-495)
            is_root = tail_positive != 0;
        } else {
            if (tail_positive) acc = f.mul(acc, f.pow_u64(x, tail_exp));
            is_root = f.is_zero(acc);
        }
        is_root = is_root && live;
        const unsigned int ballot = __ballot_sync(0xffffffffu, is_root);
        if (!WRITE) {
            if (lane == 0) my_count += __popc(ballot);
        } else {
            if (lane == 0) s_warp[warp] = __popc(ballot);
            __syncthreads();
            unsigned int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < kRootsThreads / 32; ++w) {
                const unsigned int c = s_warp[w];
                before += w < warp ? c : 0u;
                total += c;
            }
            if (is_root) {
                const unsigned long long idx = s_pos + before + __popc(ballot & ((1u << lane) - 1u));
                if (idx < max_roots) roots[idx] = (T)xi;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_pos += total;
        }
    }
    if (!WRITE) {
        if (lane == 0) s_warp[warp] = my_count;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int total = 0;
            for (int w = 0; w < kRootsThreads / 32; ++w) total += s_warp[w];
            counts[tile] = total;
        }
    }
}

// exclusive scan of counts[0..n) into offsets, total into *n_roots; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_roots_scan(const unsigned int* __restrict__ counts, size_t n,
                                                     unsigned long long* __restrict__ offsets,
                                                     unsigned long long* __restrict__ n_roots) {
    __shared__ unsigned long long s_part[1024];
    const size_t per = (n + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per;
    const size_t hi = lo + per < n ? lo + per : n;
    unsigned long long sum = 0;
    for (size_t i = lo; i < hi; ++i) sum += counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partial sums
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned long long v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0ull;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - sum;  // exclusive prefix of this thread's segment
    for (size_t i = lo; i < hi; ++i) {
        offsets[i] = run;
        run += counts[i];
    }
    if (threadIdx.x == 1023) *n_roots = s_part[1023];
}

static size_t roots_tiles(unsigned long long order) { return (size_t)((order + kRootsTile - 1) / kRootsTile); }

}  // namespace gfb

extern "C" {

size_t gfb_poly_roots_workspace_bytes(const gfb_field* f, size_t n_terms) {
    if (!f || f->order_lo == 0) return 0;
    const size_t tiles = gfb::roots_tiles(f->order_lo);
    // tile counts (u32) + tile offsets (u64) + gaps (u64 per term), each 16-byte aligned
    return ((tiles * 4 + 15) & ~(size_t)15) + tiles * 8 + n_terms * 8 + 64;
}

gfb_status gfb_poly_roots(gfb_field* f, int eb, const int64_t* h_degrees, const void* d_coeffs, size_t n_terms,
                          void* d_roots, size_t max_roots, unsigned long long* d_n_roots, void* d_workspace,
                          void* stream) {
    using namespace gfb;
    if (!f || !h_degrees || !d_coeffs || n_terms == 0 || !d_n_roots || !d_workspace) return GFB_ERR_BAD_ARG;
    if (max_roots && !d_roots) return GFB_ERR_BAD_ARG;
    // exhaustive search: the field must be enumerable (the reference's loop has the same practical bound)
    if (f->impl == IMPL_BIN_128 || f->order_lo == 0 || f->order_lo > (1ull << 33)) return GFB_ERR_UNSUPPORTED;
    for (size_t j = 0; j < n_terms; ++j) {
        if (h_degrees[j] < 0) return GFB_ERR_BAD_ARG;
        if (j && h_degrees[j] >= h_degrees[j - 1]) return GFB_ERR_BAD_ARG;  // strictly descending
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long order = f->order_lo;
    const size_t tiles = roots_tiles(order);
    char* ws = (char*)d_workspace;
    unsigned int* counts = (unsigned int*)ws;
    unsigned long long* offsets = (unsigned long long*)(ws + ((tiles * 4 + 15) & ~(size_t)15));
    unsigned long long* gaps = offsets + tiles;

    // x^gap for x != 0 only needs gap mod (order - 1); gap_0 = 0 makes the first Horner step acc = c_0.
    // A positive gap that reduces to 0 is x^(order-1) = 1: pow(x, 0) = 1 is the right value.
    const unsigned long long qm1 = order - 1;
    unsigned long long* h_gaps = (unsigned long long*)malloc(n_terms * sizeof(unsigned long long));
    if (!h_gaps) return GFB_ERR_BAD_ARG;
    h_gaps[0] = 0;
    for (size_t j = 1; j < n_terms; ++j) {
        const unsigned long long g = (unsigned long long)(h_degrees[j - 1] - h_degrees[j]);
        h_gaps[j] = g % qm1;  // order >= 2, so qm1 >= 1; a reduced gap of 1 takes the plain-multiply path (x^g = x)
    }
    bool dense = true;  // consecutive degrees: the Horner step is a plain multiply by x (gap_0 = 0 only seeds acc = c_0)
    for (size_t j = 1; j < n_terms; ++j) dense = dense && (h_degrees[j - 1] - h_degrees[j] == 1);
    const unsigned long long tail = (unsigned long long)h_degrees[n_terms - 1];
    const unsigned long long tail_exp = qm1 ? tail % qm1 : 0;
    cudaError_t e = cudaMemcpyAsync(gaps, h_gaps, n_terms * sizeof(unsigned long long), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // h_gaps is pageable: the copy is staged, but be explicit
    free(h_gaps);
    if (e != cudaSuccess) {
        set_last_error("gfb_poly_roots gaps", e);
        return GFB_ERR_CUDA;
    }

    gfb_status status = dispatch_field(f, eb, [&](auto functor, auto tag) -> gfb_status {
        using F = decltype(functor);
        using T = typename decltype(tag)::type;
        using E = typename F::elem;
        const size_t fs = field_smem_bytes(functor);
        const size_t smem = fs + kRootsChunk * (sizeof(unsigned long long) + sizeof(E));
        // grid.x is limited to 2^31 - 1 tiles: fine up to order 2^33 (2^21 tiles)
        auto launch = [&](auto dense_tag) {
            constexpr bool DENSE = decltype(dense_tag)::value;
            k_roots<T, F, false, DENSE><<<(unsigned)tiles, kRootsThreads, smem, st>>>(
                functor, (const T*)d_coeffs, gaps, n_terms, tail_exp, tail > 0, order, 0ull, counts, offsets, (T*)d_roots,
                (unsigned long long)max_roots, fs);
            k_roots_scan<<<1, 1024, 0, st>>>(counts, tiles, offsets, d_n_roots);
            k_roots<T, F, true, DENSE><<<(unsigned)tiles, kRootsThreads, smem, st>>>(
                functor, (const T*)d_coeffs, gaps, n_terms, tail_exp, tail > 0, order, 0ull, counts, offsets, (T*)d_roots,
                (unsigned long long)max_roots, fs);
        };
        if (dense)
            launch(std::true_type{});
        else
            launch(std::false_type{});
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    });
    return status;
}

gfb_status gfb_poly_roots_host(gfb_field* f, int eb, const int64_t* degrees, const void* coeffs, size_t n_terms,
                               void* roots, size_t max_roots, size_t* n_roots) {
    using namespace gfb;
    if (!f || !degrees || !coeffs || n_terms == 0 || !n_roots) return GFB_ERR_BAD_ARG;
    if (max_roots && !roots) return GFB_ERR_BAD_ARG;
    *n_roots = 0;
    if (f->impl == IMPL_BIN_128 || f->order_lo == 0 || f->order_lo > (1ull << 33)) return GFB_ERR_UNSUPPORTED;
    const size_t ws_bytes = gfb_poly_roots_workspace_bytes(f, n_terms);
    char *dc = nullptr, *dr = nullptr, *dw = nullptr;
    unsigned long long* dn = nullptr;
    cudaError_t e = cudaMalloc((void**)&dc, n_terms * (size_t)eb + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dr, max_roots * (size_t)eb + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dw, ws_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dn, sizeof(unsigned long long));
    gfb_status st = GFB_OK;
    unsigned long long count = 0;
    if (e == cudaSuccess) e = cudaMemcpy(dc, coeffs, n_terms * (size_t)eb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = gfb_poly_roots(f, eb, degrees, dc, n_terms, dr, max_roots, dn, dw, nullptr);
        if (st == GFB_OK) e = cudaMemcpy(&count, dn, sizeof(count), cudaMemcpyDeviceToHost);
        if (st == GFB_OK && e == cudaSuccess) {
            const size_t take = count < max_roots ? (size_t)count : max_roots;
            if (take) e = cudaMemcpy(roots, dr, take * (size_t)eb, cudaMemcpyDeviceToHost);
        }
    }
    cudaFree(dc);
    cudaFree(dr);
    cudaFree(dw);
    cudaFree(dn);
    if (e != cudaSuccess) {
        set_last_error("gfb_poly_roots_host", e);
        return GFB_ERR_CUDA;
    }
    *n_roots = (size_t)count;
    return st;
}

}  // extern "C"
