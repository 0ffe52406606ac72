// gfb_solve.cu — Gaussian elimination over GF(p) / GF(2^m): reduced row echelon form and (P)LU decomposition.
//
// gfb_row_reduce     replaces row_reduce_jit.__call__   (/root/reference/src/galois/_domains/_linalg.py:321-353), the
//                    engine of inv_jit (:492-522), solve_jit (:525-546), matrix_rank_jit (:479-489) and of
//                    FieldArray.row_space / column_space / left_null_space / null_space (_fields/_array.py:1541-1755).
// gfb_plu_decompose  replaces plu_decompose_jit (:390-428) and lu_decompose_jit (:356-387), which det_jit (:446-471) uses.
//
// The reference's pivoting rule is kept exactly — first non-zero entry of the column at or below the pivot row, swap,
// scale, eliminate — because with `ncols` < cols (the [A | I] forms) the rows that reduce to zero on the left carry a
// right-hand part that depends on that rule (left_null_space reads it).
//
// One column = two launches, enqueued back to back with no host round trip; the pivot row number, the skip flag and
// the running determinant live in device memory:
//   k_*_pivot     (1 CTA)   find the pivot, swap, scale; publish the pivot row (rowbuf) and the column multipliers (colbuf)
//   k_eliminate   (grid)    A[r][c] -= colbuf[r] * rowbuf[c] for every other row, columns c >= j only (the pivot row is
//                           zero to the left of column j), 256 columns x 32 rows per CTA
// Algorithmic bytes per column step: 2 * rows * (cols - j) * elem_bytes (read + write of the trailing block).
#include <climits>

#include "gfb_common.cuh"
#include "gfb_dispatch.cuh"

namespace gfb {

struct ElimState {
    int p;          // next pivot row (row_reduce)
    int cur_p;      // pivot row of the step in flight
    int skip;       // this column has no pivot: k_eliminate returns immediately
    int n_swaps;    // row exchanges so far
    int lu_failed;  // LU without pivoting met a zero pivot with a non-zero entry below it
    int pad;
    unsigned long long det;  // product of the pivots
};

constexpr int kPivotThreads = 1024;
constexpr int kElimCols = 256;
constexpr int kElimRows = 32;

template <typename F>
__global__ void k_elim_init(F f, ElimState* st, size_t fs) {
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    if (threadIdx.x == 0) {
        st->p = 0;
        st->cur_p = 0;
        st->skip = 1;
        st->n_swaps = 0;
        st->lu_failed = 0;
        st->det = (unsigned long long)f.one();
    }
}

// first row index >= lo whose entry in column j is non-zero (INT_MAX if none); all threads get the result
template <typename T>
__device__ __forceinline__ int first_nonzero(const T* __restrict__ A, size_t rows, size_t cols, size_t j, int lo,
                                             int* s_min) {
    if (threadIdx.x == 0) *s_min = INT_MAX;
    __syncthreads();
    for (size_t i = (size_t)lo + threadIdx.x; i < rows; i += blockDim.x) {
        if (A[i * cols + j] != (T)0) {
            atomicMin(s_min, (int)i);
            break;  // later rows of this thread are larger
        }
    }
    __syncthreads();
    return *s_min;
}

// ---------------------------------------------------------------------------------------- row_reduce: pivot step
template <typename T, typename F>
__global__ void __launch_bounds__(kPivotThreads)
k_rr_pivot(F f, T* __restrict__ A, size_t rows, size_t cols, size_t j, ElimState* st, typename F::elem* rowbuf,
           typename F::elem* colbuf, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    __shared__ int s_min;
    const int p = st->p;
    if ((size_t)p >= rows) {  // every row already holds a pivot (`if p == A_rre.shape[0]: break`)
        __syncthreads();
        if (threadIdx.x == 0) st->skip = 1;
        return;
    }
    const int i = first_nonzero(A, rows, cols, j, p, &s_min);
    if (i == INT_MAX) {  // `if idxs.size == 0: continue`
        if (threadIdx.x == 0) st->skip = 1;
        return;
    }
    const E piv = to_elem<E, T>(A[(size_t)i * cols + j]);
    const E pinv = f.inv1(piv);
    __syncthreads();  // everyone has read the pivot before row i is overwritten
    // swap rows p and i, then row p /= pivot
    for (size_t c = threadIdx.x; c < cols; c += blockDim.x) {
        const T vi = A[(size_t)i * cols + c];
        const T vp = A[(size_t)p * cols + c];
        const E nv = f.mul(to_elem<E, T>(vi), pinv);
        rowbuf[c] = nv;
        A[(size_t)p * cols + c] = to_store<T, E>(nv);
        if (i != p) A[(size_t)i * cols + c] = vp;
    }
    __syncthreads();
    // multipliers of every other row: their entry in column j (after the swap)
    for (size_t r = threadIdx.x; r < rows; r += blockDim.x)
        colbuf[r] = r == (size_t)p ? f.zero() : to_elem<E, T>(A[r * cols + j]);
    if (threadIdx.x == 0) {
        st->cur_p = p;
        st->p = p + 1;
        st->skip = 0;
        st->det = (unsigned long long)f.mul((E)st->det, piv);
        if (i != p) st->n_swaps += 1;
    }
}

// ---------------------------------------------------------------------------------------- (P)LU: pivot step
template <typename T, typename F>
__global__ void __launch_bounds__(kPivotThreads)
k_lu_pivot(F f, T* __restrict__ A, T* __restrict__ L, long long* __restrict__ perm, size_t rows, size_t cols, size_t i,
           int pivoting, ElimState* st, typename F::elem* rowbuf, typename F::elem* colbuf, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    __shared__ int s_min;
    if (A[i * cols + i] == (T)0) {  // uniform: every thread reads the same word
        const int r = first_nonzero(A, rows, cols, i, (int)i, &s_min);
        if (r == INT_MAX) {  // the whole column is zero from the diagonal down: L[i, i] = 1, next column
            if (threadIdx.x == 0) {
                L[i * rows + i] = (T)1;
                st->skip = 1;
            }
            return;
        }
        if (!pivoting) {  // "The LU decomposition of 'A' does not exist."
            if (threadIdx.x == 0) {
                st->lu_failed = 1;
                st->skip = 1;
            }
            return;
        }
        // swap rows i and r of Ai, of L and of the permutation
        for (size_t c = threadIdx.x; c < cols; c += blockDim.x) {
            const T a = A[i * cols + c];
            A[i * cols + c] = A[(size_t)r * cols + c];
            A[(size_t)r * cols + c] = a;
        }
        for (size_t c = threadIdx.x; c < rows; c += blockDim.x) {
            const T a = L[i * rows + c];
            L[i * rows + c] = L[(size_t)r * rows + c];
            L[(size_t)r * rows + c] = a;
        }
        if (threadIdx.x == 0) {
            const long long t = perm[i];
            perm[i] = perm[r];
            perm[r] = t;
            st->n_swaps += 1;
        }
        __syncthreads();
    }
    const E piv = to_elem<E, T>(A[i * cols + i]);
    const E pinv = f.inv1(piv);
    for (size_t c = threadIdx.x; c < cols; c += blockDim.x) rowbuf[c] = to_elem<E, T>(A[i * cols + c]);
    // l = Ai[i+1:, i] / Ai[i, i];  L[i+1:, i] = l
    for (size_t r = threadIdx.x; r < rows; r += blockDim.x) {
        E l = f.zero();
        if (r > i) {
            l = f.mul(to_elem<E, T>(A[r * cols + i]), pinv);
            L[r * rows + i] = to_store<T, E>(l);
        }
        colbuf[r] = l;
    }
    if (threadIdx.x == 0) {
        L[i * rows + i] = (T)1;
        st->cur_p = (int)i;
        st->skip = 0;
    }
}

// ---------------------------------------------------------------------------------------- shared elimination step
template <typename T, typename F>
__global__ void __launch_bounds__(kElimCols)
k_eliminate(F f, T* __restrict__ A, size_t rows, size_t cols, size_t c0, const ElimState* __restrict__ st,
            const typename F::elem* __restrict__ rowbuf, const typename F::elem* __restrict__ colbuf, size_t fs) {
    using E = typename F::elem;
    if (st->skip) return;  // uniform across the grid
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    const size_t c = c0 + (size_t)blockIdx.x * kElimCols + threadIdx.x;
    if (c >= cols) return;
    const E pr = rowbuf[c];
    if (f.is_zero(pr)) return;  // nothing to subtract in this column
    const size_t r0 = (size_t)blockIdx.y * kElimRows;
    const size_t r1 = r0 + kElimRows < rows ? r0 + kElimRows : rows;
    for (size_t r = r0; r < r1; ++r) {
        const E fac = colbuf[r];  // 0 for the pivot row itself (and the rows above it in the LU forms)
        if (f.is_zero(fac)) continue;
        const size_t idx = r * cols + c;
        A[idx] = to_store<T, E>(f.sub(to_elem<E, T>(A[idx]), f.mul(fac, pr)));
    }
}

template <typename T>
__global__ void k_lu_finish(T* L, long long* perm_unused, size_t rows) {
    (void)perm_unused;
    L[(rows - 1) * rows + (rows - 1)] = (T)1;  // `L[-1, -1] = 1`
}

template <typename T>
__global__ void k_lu_init(T* L, long long* perm, size_t rows, int identity) {
    const size_t total = rows * rows;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / rows, c = i - r * rows;
        L[i] = (identity && r == c) ? (T)1 : (T)0;
        if (c == 0) perm[r] = (long long)r;
    }
}

template <typename T, typename F>
__global__ void k_elim_report(F f, const ElimState* st, long long* info, T* det, int mode, size_t fs) {
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    using E = typename F::elem;
    if (threadIdx.x != 0) return;
    if (mode == 0) {  // row_reduce
        info[0] = st->p;
        info[1] = st->n_swaps;
        if (det) {
            E d = (E)st->det;
            if (st->n_swaps & 1) d = f.neg(d);
            *det = to_store<T, E>(d);
        }
    } else {  // (P)LU
        info[0] = st->n_swaps;
        info[1] = st->lu_failed;
    }
}

struct ElimWorkspace {
    ElimState* state;
    void* rowbuf;
    void* colbuf;
};

static ElimWorkspace carve(void* ws, size_t cols) {
    char* base = (char*)ws;
    ElimWorkspace w;
    w.state = (ElimState*)base;
    w.rowbuf = base + 64;
    w.colbuf = base + 64 + ((cols * 8 + 15) & ~(size_t)15);
    return w;
}

static dim3 elim_grid(size_t rows, size_t cols, size_t c0) {
    return dim3((unsigned)((cols - c0 + kElimCols - 1) / kElimCols), (unsigned)((rows + kElimRows - 1) / kElimRows));
}

}  // namespace gfb

extern "C" {

size_t gfb_row_reduce_workspace_bytes(const gfb_field* f, size_t rows, size_t cols) {
    (void)f;
    return 64 + ((cols * 8 + 15) & ~(size_t)15) + rows * 8 + 64;
}

gfb_status gfb_row_reduce(gfb_field* f, int eb, void* A, size_t rows, size_t cols, size_t ncols, int64_t* d_info,
                          void* d_det, void* d_workspace, void* stream) {
    using namespace gfb;
    if (!f || !d_info || !d_workspace) return GFB_ERR_BAD_ARG;
    if (ncols > cols) return GFB_ERR_BAD_ARG;
    if (rows > (size_t)INT_MAX / 2 || cols > (size_t)INT_MAX / 2) return GFB_ERR_UNSUPPORTED;
    if ((rows + kElimRows - 1) / kElimRows > 65535) return GFB_ERR_UNSUPPORTED;
    if (rows && cols && !A) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_field(f, eb, [&](auto functor, auto tag) -> gfb_status {
        using F = decltype(functor);
        using T = typename decltype(tag)::type;
        using E = typename F::elem;
        const size_t fs = field_smem_bytes(functor);
        ElimWorkspace w = carve(d_workspace, cols);
        k_elim_init<F><<<1, 32, fs, st>>>(functor, w.state, fs);
        if (rows && cols) {
            for (size_t j = 0; j < ncols; ++j) {
                k_rr_pivot<T, F><<<1, kPivotThreads, fs, st>>>(functor, (T*)A, rows, cols, j, w.state, (E*)w.rowbuf,
                                                                (E*)w.colbuf, fs);
                k_eliminate<T, F><<<elim_grid(rows, cols, j), kElimCols, fs, st>>>(
                    functor, (T*)A, rows, cols, j, w.state, (const E*)w.rowbuf, (const E*)w.colbuf, fs);
            }
        }
        k_elim_report<T, F><<<1, 32, fs, st>>>(functor, w.state, (long long*)d_info, (T*)d_det, 0, fs);
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    });
}

gfb_status gfb_plu_decompose(gfb_field* f, int eb, void* A, void* L, int64_t* d_perm, size_t rows, size_t cols,
                             int pivoting, int64_t* d_info, void* d_workspace, void* stream) {
    using namespace gfb;
    if (!f || !A || !L || !d_perm || !d_info || !d_workspace || rows == 0 || cols == 0) return GFB_ERR_BAD_ARG;
    if (rows > (size_t)INT_MAX / 2 || cols > (size_t)INT_MAX / 2) return GFB_ERR_UNSUPPORTED;
    if ((rows + kElimRows - 1) / kElimRows > 65535) return GFB_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_field(f, eb, [&](auto functor, auto tag) -> gfb_status {
        using F = decltype(functor);
        using T = typename decltype(tag)::type;
        using E = typename F::elem;
        const size_t fs = field_smem_bytes(functor);
        ElimWorkspace w = carve(d_workspace, cols);
        k_elim_init<F><<<1, 32, fs, st>>>(functor, w.state, fs);
        size_t blocks = (rows * rows + 255) / 256;
        if (blocks > 4096) blocks = 4096;
        // plu: L starts as zeros and gains its unit diagonal step by step; lu: L starts as the identity
        k_lu_init<T><<<(unsigned)blocks, 256, 0, st>>>((T*)L, (long long*)d_perm, rows, pivoting ? 0 : 1);
        // plu: range(0, min(m, n)) (:404); lu: range(0, m - 1) (:371), which the caller has checked fits the columns
        const size_t steps = pivoting ? (rows < cols ? rows : cols) : rows - 1;
        for (size_t i = 0; i < steps; ++i) {
            k_lu_pivot<T, F><<<1, kPivotThreads, fs, st>>>(functor, (T*)A, (T*)L, (long long*)d_perm, rows, cols, i,
                                                            pivoting, w.state, (E*)w.rowbuf, (E*)w.colbuf, fs);
            k_eliminate<T, F><<<elim_grid(rows, cols, i), kElimCols, fs, st>>>(
                functor, (T*)A, rows, cols, i, w.state, (const E*)w.rowbuf, (const E*)w.colbuf, fs);
        }
        if (pivoting) k_lu_finish<T><<<1, 1, 0, st>>>((T*)L, nullptr, rows);
        k_elim_report<T, F><<<1, 32, fs, st>>>(functor, w.state, (long long*)d_info, (T*)nullptr, 1, fs);
        GFB_CUDA_TRY(cudaGetLastError());
        return GFB_OK;
    });
}

gfb_status gfb_row_reduce_host(gfb_field* f, int eb, const void* A, void* A_rre, size_t rows, size_t cols, size_t ncols,
                               int64_t* n_pivots, void* det) {
    using namespace gfb;
    if (!f || !n_pivots) return GFB_ERR_BAD_ARG;
    if (rows && cols && (!A || !A_rre)) return GFB_ERR_BAD_ARG;
    *n_pivots = 0;
    const size_t bytes = rows * cols * (size_t)eb;
    char *dA = nullptr, *dw = nullptr, *dd = nullptr;
    long long* di = nullptr;
    cudaError_t e = cudaMalloc((void**)&dA, bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dw, gfb_row_reduce_workspace_bytes(f, rows, cols));
    if (e == cudaSuccess) e = cudaMalloc((void**)&di, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dd, 16);
    gfb_status st = GFB_OK;
    if (e == cudaSuccess && bytes) e = cudaMemcpy(dA, A, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = gfb_row_reduce(f, eb, dA, rows, cols, ncols, (int64_t*)di, det ? dd : nullptr, dw, nullptr);
        long long info[2] = {0, 0};
        if (st == GFB_OK) e = cudaMemcpy(info, di, sizeof(info), cudaMemcpyDeviceToHost);
        if (st == GFB_OK && e == cudaSuccess && bytes) e = cudaMemcpy(A_rre, dA, bytes, cudaMemcpyDeviceToHost);
        if (st == GFB_OK && e == cudaSuccess && det) e = cudaMemcpy(det, dd, (size_t)eb, cudaMemcpyDeviceToHost);
        *n_pivots = info[0];
    }
    cudaFree(dA);
    cudaFree(dw);
    cudaFree(di);
    cudaFree(dd);
    if (e != cudaSuccess) {
        set_last_error("gfb_row_reduce_host", e);
        return GFB_ERR_CUDA;
    }
    return st;
}

gfb_status gfb_plu_decompose_host(gfb_field* f, int eb, const void* A, void* U, void* L, int64_t* perm, size_t rows,
                                  size_t cols, int pivoting, int64_t* n_swaps, int* failed) {
    using namespace gfb;
    if (!f || !A || !U || !L || !perm || !n_swaps || !failed || rows == 0 || cols == 0) return GFB_ERR_BAD_ARG;
    const size_t bytes = rows * cols * (size_t)eb, lbytes = rows * rows * (size_t)eb;
    char *dA = nullptr, *dL = nullptr, *dw = nullptr;
    long long *di = nullptr, *dp = nullptr;
    cudaError_t e = cudaMalloc((void**)&dA, bytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dL, lbytes + 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dw, gfb_row_reduce_workspace_bytes(f, rows, cols));
    if (e == cudaSuccess) e = cudaMalloc((void**)&di, 4 * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc((void**)&dp, rows * sizeof(long long));
    gfb_status st = GFB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(dA, A, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        st = gfb_plu_decompose(f, eb, dA, dL, (int64_t*)dp, rows, cols, pivoting, (int64_t*)di, dw, nullptr);
        long long info[2] = {0, 0};
        if (st == GFB_OK) e = cudaMemcpy(info, di, sizeof(info), cudaMemcpyDeviceToHost);
        if (st == GFB_OK && e == cudaSuccess) e = cudaMemcpy(U, dA, bytes, cudaMemcpyDeviceToHost);
        if (st == GFB_OK && e == cudaSuccess) e = cudaMemcpy(L, dL, lbytes, cudaMemcpyDeviceToHost);
        if (st == GFB_OK && e == cudaSuccess) e = cudaMemcpy(perm, dp, rows * sizeof(long long), cudaMemcpyDeviceToHost);
        *n_swaps = info[0];
        *failed = (int)info[1];
    }
    cudaFree(dA);
    cudaFree(dL);
    cudaFree(dw);
    cudaFree(di);
    cudaFree(dp);
    if (e != cudaSuccess) {
        set_last_error("gfb_plu_decompose_host", e);
        return GFB_ERR_CUDA;
    }
    return st;
}

}  // extern "C"
