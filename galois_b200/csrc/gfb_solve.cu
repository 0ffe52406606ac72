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
// row_reduce is BLOCKED: 32 pivot columns (a panel) per pass over the matrix, everything enqueued with no host round
// trip (pivot row counter, exchange count and running determinant live in device memory):
//   k_rr_panel         (1 CTA)  the 32 column steps restricted to the rows x 32 slab: pivot search, exchange, scale,
//                               eliminate — and record, per step, the pivot position, its exchange partner, the pivot
//                               inverse and the multiplier of every row (M, rows x 32; later exchanges permute it too)
//   k_rr_panel_w       (grid)   per trailing column: replay the exchanges, solve the 32 x 32 triangular recurrence for the
//                               scaled pivot-row entries w_s, finish the pivot rows, publish W (32 x cols)
//   k_rr_panel_update  (grid)   A[r][c] -= sum_s M[r][s] * W[s][c] for every non-pivot row: ONE read + write of the
//                               trailing block per 32 columns instead of per column
// Exact field arithmetic makes the blocked order bit-identical to the reference's column-by-column order (the pivot
// decisions only ever look at fully updated slab columns).  Columns left of the panel are never touched: the pivot rows
// are zero there.  Algorithmic bytes per panel: 2 * rows * (cols - j0 - 32) * elem_bytes.
//
// (P)LU keeps the simple form, two launches per column:
//   k_lu_pivot    (1 CTA)   find the pivot, exchange; publish the pivot row (rowbuf) and the column multipliers (colbuf)
//   k_eliminate   (grid)    A[r][c] -= colbuf[r] * rowbuf[c] for the rows below, columns c >= i, 256 columns x 32 rows per CTA
#include <climits>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include <cooperative_groups.h>

#include "gfb_common.cuh"
#include "gfb_dispatch.cuh"
#include "gfb_dot.cuh"

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

// ---------------------------------------------------------------------------------------- row_reduce: blocked form
constexpr int kPanel = 32;

struct PanelMeta {
    int piv_p[kPanel];                // pivot position of step s, -1 when column j0 + s has no pivot
    int piv_i[kPanel];                // the row it was exchanged with (== piv_p when no exchange)
    unsigned long long pinv[kPanel];  // inverse of the pivot
    int p0, p1;                       // this panel's pivot rows are [p0, p1)
    int pad[2];
};

// The slab (rows x kw entries of the panel columns) is staged in shared memory when it fits (SMEM = true, row stride
// kSlabStride = 33 words: the column scans and the per-row updates are both bank-conflict free); otherwise it stays in
// global memory and each row segment is updated through fully unrolled, independent loads.
constexpr int kSlabStride = kPanel + 1;

// storage type of the panel slab: in shared memory a field whose elements fit 32 bits is held as uint32 even when the
// array dtype is wider (a 1024-row slab of an int64 array then fits beside the field tables)
template <typename T, typename E, bool SMEM>
using SlabT = std::conditional_t<(SMEM && sizeof(E) < sizeof(T)), E, T>;

template <typename T, typename F, bool SMEM>
__global__ void __launch_bounds__(kPivotThreads)
k_rr_panel(F f, T* __restrict__ A, size_t rows, size_t cols, size_t j0, int kw, ElimState* st, PanelMeta* meta,
           typename F::elem* __restrict__ M, size_t fs) {
    using E = typename F::elem;
    using S = SlabT<T, E, SMEM>;  // the shared-memory slab holds field elements in their compute width
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    S* s_slab = reinterpret_cast<S*>(reinterpret_cast<char*>(smem_raw) + fs);
    __shared__ int s_min;
    __shared__ E s_prow[kPanel];
    __shared__ E s_pinv;
    auto at = [&](size_t r, int t) -> S& {
        if constexpr (SMEM)
            return s_slab[r * kSlabStride + t];
        else
            return A[r * cols + j0 + t];
    };
    if constexpr (SMEM) {
        for (size_t idx = threadIdx.x; idx < rows * (size_t)kw; idx += blockDim.x) {
            const size_t r = idx / kw;
            const int t = (int)(idx - r * kw);
            s_slab[r * kSlabStride + t] = (S)A[r * cols + j0 + t];
        }
    }
    int p = st->p;
    const int p0 = p;
    E det = (E)st->det;
    int swaps = 0;
    __syncthreads();
    for (int s = 0; s < kw; ++s) {
        // first non-zero entry of column s of the slab at or below row p (p is uniform across the CTA)
        int i = INT_MAX;
        if ((size_t)p < rows) {
            if (threadIdx.x == 0) s_min = INT_MAX;
            __syncthreads();
            int cand = INT_MAX;
            for (size_t r = (size_t)p + threadIdx.x; r < rows; r += blockDim.x) {
                if (at(r, s) != (S)0) {
                    cand = (int)r;
                    break;  // later rows of this thread are larger
                }
            }
            cand = __reduce_min_sync(0xffffffffu, cand);  // one shared-memory atomic per warp, not per row
            if ((threadIdx.x & 31) == 0 && cand != INT_MAX) atomicMin(&s_min, cand);
            __syncthreads();
            i = s_min;
        }
        if (i == INT_MAX) {  // no pivot in this column (or every row already has one): a null step
            for (size_t r = threadIdx.x; r < rows; r += blockDim.x) M[r * kPanel + s] = f.zero();
            if (threadIdx.x == 0) meta->piv_p[s] = -1;
            __syncthreads();  // s_min is rewritten by the next step
            continue;
        }
        const E piv = to_elem<E, S>(at(i, s));
        __syncthreads();  // the pivot has been read by everyone before row i moves
        // thread 0 inverts the pivot (a ~45-multiply dependent chain) while warps 1 and 2 exchange rows p and i of the
        // slab and of the multipliers recorded by the earlier steps
        if (threadIdx.x == 0) {
            s_pinv = f.inv1(piv);
        } else if (i != p && threadIdx.x >= 32 && threadIdx.x < 96) {
            const int t = (threadIdx.x & 31);
            if (threadIdx.x < 64) {
                if (t < kw) {
                    const S a = at(p, t);
                    at(p, t) = at(i, t);
                    at(i, t) = a;
                }
            } else if (t < s) {
                const E a = M[(size_t)p * kPanel + t];
                M[(size_t)p * kPanel + t] = M[(size_t)i * kPanel + t];
                M[(size_t)i * kPanel + t] = a;
            }
        }
        __syncthreads();
        const E pinv = s_pinv;
        if ((int)threadIdx.x < kw) {  // row p /= pivot (its entries left of column j0 + s are already zero)
            const E v = f.mul(to_elem<E, S>(at(p, threadIdx.x)), pinv);
            s_prow[threadIdx.x] = f.premul(v);  // multiplier form: one reduction per product below
            at(p, threadIdx.x) = to_store<S, E>(v);
        }
        __syncthreads();
        for (size_t r = threadIdx.x; r < rows; r += blockDim.x) {
            E m = f.zero();
            if (r != (size_t)p) {
                m = to_elem<E, S>(at(r, s));
                if (!f.is_zero(m)) {
#pragma unroll 4
                    for (int t = s; t < kw; ++t)
                        at(r, t) = to_store<S, E>(f.sub(to_elem<E, S>(at(r, t)), f.mul_prepared(m, s_prow[t])));
                }
            }
            M[r * kPanel + s] = m;
        }
        if (threadIdx.x == 0) {
            meta->piv_p[s] = p;
            meta->piv_i[s] = i;
            meta->pinv[s] = (unsigned long long)pinv;
        }
        det = f.mul(det, piv);
        swaps += i != p ? 1 : 0;
        ++p;
        __syncthreads();
    }
    if constexpr (SMEM) {
        for (size_t idx = threadIdx.x; idx < rows * (size_t)kw; idx += blockDim.x) {
            const size_t r = idx / kw;
            const int t = (int)(idx - r * kw);
            A[r * cols + j0 + t] = (T)s_slab[r * kSlabStride + t];
        }
    }
    if (threadIdx.x == 0) {
        st->p = p;
        st->det = (unsigned long long)det;
        st->n_swaps += swaps;
        meta->p0 = p0;
        meta->p1 = p;
    }
}

constexpr int kPanelWThreads = 128;

template <typename T, typename F>
__global__ void __launch_bounds__(kPanelWThreads)
k_rr_panel_w(F f, T* __restrict__ A, size_t rows, size_t cols, size_t c0, int kw, const PanelMeta* __restrict__ meta,
             const typename F::elem* __restrict__ M, typename F::elem* __restrict__ W, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    (void)rows;
    __shared__ E s_mp[kPanel][kPanel + 1];  // s_mp[s][t] = multiplier of pivot row s at step t
    __shared__ int s_p[kPanel], s_i[kPanel];
    __shared__ E s_pinv[kPanel];
    __shared__ int s_pos[2 * kPanel], s_src[2 * kPanel], s_moves;
    __shared__ E s_w[kPanel][kPanelWThreads];
    if (meta->p1 == meta->p0) return;  // the panel found no pivot: nothing changes (uniform)
    for (int idx = threadIdx.x; idx < kPanel * kPanel; idx += blockDim.x) {
        const int s = idx / kPanel, t = idx - s * kPanel;
        const int ps = s < kw ? meta->piv_p[s] : -1;
        s_mp[s][t] = (ps >= 0 && t < kw) ? M[(size_t)ps * kPanel + t] : f.zero();
    }
    if (threadIdx.x < kPanel) {
        const int s = threadIdx.x;
        s_p[s] = s < kw ? meta->piv_p[s] : -1;
        s_i[s] = s < kw ? meta->piv_i[s] : -1;
        s_pinv[s] = s < kw ? (E)meta->pinv[s] : f.zero();
    }
    __syncthreads();
    // Net effect of the panel's row exchanges as a move list (built once per CTA from the shared copies of the step
    // records): position s_pos[k] ends up holding the row that started at s_src[k].  Every thread then moves its own
    // column with independent loads and stores instead of replaying up to 32 dependent exchanges.
    if (threadIdx.x == 0) {
        int n = 0;
        for (int s = 0; s < kw; ++s) {
            const int ps = s_p[s], is = s_i[s];
            if (ps < 0 || is == ps) continue;
            int kp = -1, ki = -1;
            for (int k = 0; k < n; ++k) {
                if (s_pos[k] == ps) kp = k;
                if (s_pos[k] == is) ki = k;
            }
            if (kp < 0) {
                kp = n++;
                s_pos[kp] = ps;
                s_src[kp] = ps;
            }
            if (ki < 0) {
                ki = n++;
                s_pos[ki] = is;
                s_src[ki] = is;
            }
            const int a = s_src[kp];
            s_src[kp] = s_src[ki];
            s_src[ki] = a;
        }
        s_moves = n;
    }
    __syncthreads();
    const size_t c = c0 + (size_t)blockIdx.x * kPanelWThreads + threadIdx.x;
    if (c >= cols) return;
    {
        const int n = s_moves;
        T v[2 * kPanel];
        for (int k = 0; k < n; ++k) v[k] = A[(size_t)s_src[k] * cols + c];
        for (int k = 0; k < n; ++k)
            if (s_src[k] != s_pos[k]) A[(size_t)s_pos[k] * cols + c] = v[k];
    }
    // w_s = pinv_s * (x[p_s] - sum_{t<s} M[p_s][t] w_t): the entry of the scaled pivot row at the moment step s used it.
    // Rolled loops over a per-thread column of shared memory (s_w[.][tid]): fully unrolling the 32 x 32 recurrence makes
    // ~230 KB of straight-line code that runs at instruction-fetch speed.
    const int tid = threadIdx.x;
#pragma unroll 8
    for (int s = 0; s < kw; ++s) s_w[s][tid] = to_elem<E, T>(A[(size_t)(s_p[s] >= 0 ? s_p[s] : 0) * cols + c]);
#pragma unroll 1
    for (int s = 0; s < kw; ++s) {
        E ws = f.zero();
        if (s_p[s] >= 0) {
            DotAcc<F> acc;
            acc.init(f);
#pragma unroll 4
            for (int t = 0; t < s; ++t) acc.fma(f, s_mp[s][t], s_w[t][tid]);
            ws = f.mul(s_pinv[s], f.sub(s_w[s][tid], acc.done(f)));
        }
        s_w[s][tid] = ws;
    }
    // the pivot rows end as w_s minus what the later steps subtracted from them
#pragma unroll 1
    for (int s = 0; s < kw; ++s) {
        if (s_p[s] >= 0) {
            DotAcc<F> acc;
            acc.init(f);
#pragma unroll 4
            for (int t = s + 1; t < kw; ++t) acc.fma(f, s_mp[s][t], s_w[t][tid]);
            A[(size_t)s_p[s] * cols + c] = to_store<T, E>(f.sub(s_w[s][tid], acc.done(f)));
        }
    }
#pragma unroll 4
    for (int s = 0; s < kw; ++s) W[(size_t)s * cols + c] = s_w[s][tid];
}

constexpr int kUpdCols = 128;
constexpr int kUpdRows = 16;

// SMALLP: a prime below 2^30 — sixteen products fit one 64-bit accumulator, so a multiply-add is ONE IMAD.WIDE instead of a
// 128-bit add-with-carry chain (the kernel is bound by the integer-multiply pipe: ncu, profiles/r02_summary.md).
template <typename T, typename F, bool SMALLP = false>
__global__ void __launch_bounds__(kUpdCols)
k_rr_panel_update(F f, T* __restrict__ A, size_t rows, size_t cols, size_t c0, size_t c_end, int kw,
                  const PanelMeta* __restrict__ meta, const typename F::elem* __restrict__ M,
                  const typename F::elem* __restrict__ W, size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    __shared__ E s_m[kUpdRows][kPanel];
    const int p0 = meta->p0, p1 = meta->p1;
    if (p1 == p0) return;  // uniform
    const size_t r0 = (size_t)blockIdx.y * kUpdRows;
    for (int idx = threadIdx.x; idx < kUpdRows * kPanel; idx += blockDim.x) {
        const int rr = idx / kPanel, t = idx - rr * kPanel;
        const size_t r = r0 + rr;
        s_m[rr][t] = (r < rows && t < kw) ? M[r * kPanel + t] : f.zero();
    }
    __syncthreads();
    const size_t c = c0 + (size_t)blockIdx.x * kUpdCols + threadIdx.x;
    if (c >= c_end) return;  // columns [c0, c_end)
    E w[kPanel];
#pragma unroll
    for (int t = 0; t < kPanel; ++t) w[t] = t < kw ? W[(size_t)t * cols + c] : f.zero();
    for (int rr = 0; rr < kUpdRows; ++rr) {
        const size_t r = r0 + rr;
        if (r >= rows) break;
        if (r >= (size_t)p0 && r < (size_t)p1) continue;  // pivot rows were finished by k_rr_panel_w
        E d;
        if constexpr (SMALLP) {
            unsigned long long s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int t = 0; t < kPanel / 2; ++t) {
                s0 += (unsigned long long)s_m[rr][t] * w[t];
                s1 += (unsigned long long)s_m[rr][t + kPanel / 2] * w[t + kPanel / 2];
            }
            d = f.add(f.reduce64(s0), f.reduce64(s1));
        } else {
            DotAcc<F> acc;
            acc.init(f);
#pragma unroll
            for (int t = 0; t < kPanel; ++t) acc.fma(f, s_m[rr][t], w[t]);
            d = acc.done(f);
        }
        if (!f.is_zero(d)) {
            const size_t idx = r * cols + c;
            A[idx] = to_store<T, E>(f.sub(to_elem<E, T>(A[idx]), d));
        }
    }
}

// ---------------------------------------------------------------------------------------- row_reduce: cluster panel
// The panel step spread over a thread-block cluster (fields whose elements fit 32 bits, up to 4096 rows).  k_rr_panel above
// is one CTA walking a shared-memory slab: every column step pays a Fermat inversion of the pivot on one thread (~1 us), five
// CTA barriers and a read-modify-write pass over the slab — 4.4 us per step, 142 us per 32-column panel of a 1024-row matrix,
// two thirds of the whole elimination.  Here
//   * eight CTAs x 128 threads form one cluster; every thread keeps its row(s) of the slab IN REGISTERS (32 entries each,
//     rotated by one register per step so that the current column is always entry 0 — no dynamic register indexing);
//   * rows are never moved: each row carries its logical position, an exchange swaps two integers;
//   * the pivot search is a warp min + one distributed-shared-memory atomicMin per warp into every CTA's copy, the owner's
//     warp pushes the pivot row (32 values) into every CTA's shared memory, two cluster barriers per step;
//   * the steps are FRACTION-FREE: row_r <- piv * row_r - a[r][s] * row_p (every row but the pivot row, which is only
//     scaled), so no inversion sits on the critical path.  All rows then share the scale c(s) = prod of the earlier scaled
//     pivots, the entry a[r][s] a row holds when column s is eliminated is c(s) times the multiplier the textbook step would
//     use (times the true pivot of row r's own step when r already is a pivot row), and the 65 inversions that undo the
//     scaling (c(s), the scaled pivots, the final scale) run in parallel once per panel.  The row-reduced form is unique,
//     and the multipliers / pivots / exchange records written for the trailing kernels are exactly the textbook ones.
//   * the trailing columns do not replay the 32-step recurrence per column (k_rr_panel_w: a 55 us dependent chain on 14
//     SMs): the panel also emits the two 32 x 32 matrices T (w = T x: the scaled pivot-row entries the steps used) and
//     T2 (the final pivot rows = T2 x) and k_rr_panel_w2 applies them as 64 independent dot products per column.
namespace cgx = cooperative_groups;
constexpr int kClCtas = 8;
constexpr int kClThreads = 128;
constexpr int kClRows = kClCtas * kClThreads;

// ODDP (PrimeBarrett only): the prime is odd, products go straight to the Montgomery step.  mul_prepared() tests p == 2 on
// every call and the compiler keeps that as a (uniform) branch: one convergence region per entry of the register update, no
// overlap between the 31 independent updates of a step — with one warp per scheduler the step then runs at one dependent
// instruction per ~6 clocks (ncu: profiles/r02_elim_cluster.md).
template <typename T, typename F, int RPT, bool ODDP = false>
__global__ void __cluster_dims__(kClCtas, 1, 1) __launch_bounds__(kClThreads)
k_rr_panel_cluster(F f, T* __restrict__ A, size_t rows, size_t cols, size_t j0, int kw, ElimState* st, PanelMeta* meta,
                   typename F::elem* __restrict__ M, typename F::elem* __restrict__ TT, size_t fs) {
    using E = typename F::elem;
    static_assert(sizeof(E) == 4, "the cluster panel keeps 32-bit field elements in registers");
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    cgx::cluster_group cluster = cgx::this_cluster();
    const unsigned rank = cluster.block_rank();
    __shared__ int s_min[2];
    __shared__ E s_prow[2][kPanel];  // pivot row of the step in flight, multiplier form (every CTA has its own copy)
    __shared__ E s_piv[2];           // its (scaled) pivot
    __shared__ E s_pivs[kPanel], s_cs[kPanel], s_cinv[kPanel], s_pinvt[kPanel];
    __shared__ E s_cendinv;
    __shared__ int s_pp[kPanel], s_pi[kPanel];
    __shared__ E s_mp[kPanel][kPanel + 1];  // rank 0: s_mp[s][t] = multiplier of pivot row s at step t
    const int tid = threadIdx.x, lane = tid & 31;

    E a[RPT][kPanel];
    int pos[RPT], stp[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const size_t r = (size_t)k * kClRows + (size_t)rank * kClThreads + tid;
        stp[k] = -1;
        if (r < rows) {
            pos[k] = (int)r;
#pragma unroll
            for (int t = 0; t < kPanel; ++t) a[k][t] = t < kw ? to_elem<E, T>(A[r * cols + j0 + t]) : f.zero();
        } else {
            pos[k] = INT_MAX;
#pragma unroll
            for (int t = 0; t < kPanel; ++t) a[k][t] = f.zero();
        }
    }
    if (tid < 2) s_min[tid] = INT_MAX;
    for (int idx = tid; idx < kPanel * (kPanel + 1); idx += kClThreads) (&s_mp[0][0])[idx] = f.zero();
    int p = st->p;
    const int p0 = p;
    E c = f.one();
    uint32_t nullmask = 0;
    int swaps = 0;
    cluster.sync();  // every CTA's s_min is initialised before the first remote atomic

    // One step = search barrier, owner's push, push barrier, update.  The search of step s + 1 is software-pipelined into
    // step s: the next column (entry 1) is updated first, its candidates are sent and the search barrier is ARRIVED at
    // before the other 30 entries are updated, so that barrier's latency hides behind the update.
    auto post_candidate = [&](E (&col)[RPT], int p_now, int buf) {
        int cand = INT_MAX;
#pragma unroll
        for (int k = 0; k < RPT; ++k)
            if (!f.is_zero(col[k]) && pos[k] >= p_now && pos[k] < cand) cand = pos[k];
        cand = __reduce_min_sync(0xffffffffu, cand);
        if (lane == 0 && cand != INT_MAX) {
#pragma unroll
            for (int r = 0; r < kClCtas; ++r) atomicMin(cluster.map_shared_rank(&s_min[buf], r), cand);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    };
    if (kw > 0) {
        E col[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) col[k] = a[k][0];
        post_candidate(col, p, 0);
    }
    for (int s = 0; s < kPanel; ++s) {
        if (s < kw) {  // uniform
            const int b = s & 1;
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // the search of column s is complete
            const int i = s_min[b];
            if (tid == 0) s_min[b ^ 1] = INT_MAX;  // for step s + 1: its atomics start after the second barrier below
            if (i != INT_MAX) {
                bool own[RPT];
                bool own_any = false;
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    own[k] = pos[k] == i;
                    own_any |= own[k];
                }
                const uint32_t ow = __ballot_sync(0xffffffffu, own_any);
                if (ow) {  // the owner's warp publishes the pivot row: lane L takes column L and writes it to every CTA
                    const int ol = __ffs(ow) - 1;
                    E mine = f.zero();
#pragma unroll
                    for (int u = 0; u < kPanel; ++u) {
                        E v = f.zero();
#pragma unroll
                        for (int k = 0; k < RPT; ++k)
                            if (own[k]) v = a[k][u];
                        const E got = __shfl_sync(0xffffffffu, v, ol);
                        if (lane == ((s + u) & 31)) mine = got;
                    }
                    const E pm = f.premul(mine);
#pragma unroll
                    for (int r = 0; r < kClCtas; ++r) {
                        *cluster.map_shared_rank(&s_prow[b][lane], r) = pm;
                        if (lane == (s & 31)) *cluster.map_shared_rank(&s_piv[b], r) = mine;
                    }
                }
#pragma unroll
                for (int k = 0; k < RPT; ++k) {  // the exchange of logical rows p and i
                    if (own[k]) {
                        pos[k] = p;
                        stp[k] = s;
                    } else if (pos[k] == p) {
                        pos[k] = i;
                    }
                }
            }
            cluster.sync();
            const bool found = i != INT_MAX;
            E piv = f.zero(), pivP = f.zero();
            // entry u is column (s + u) mod 32: it changes when that is a later column or an earlier one that had no pivot
            auto mulp = [&](E x, E y_prepared) -> E {
                if constexpr (ODDP)
                    return f.montmul(x, y_prepared);
                else
                    return f.mul_prepared(x, y_prepared);
            };
            auto update_entry = [&](int k, int u, E m, bool is_piv_row, E cur) -> E {
                const bool need = u <= kPanel - 1 - s || ((nullmask >> ((s + u) & 31)) & 1u);
                (void)is_piv_row;  // the pivot row comes in with m = 0: it is only scaled (no branch inside the entry)
                const E y = f.sub(mulp(cur, pivP), mulp(m, s_prow[b][(s + u) & 31]));
                return need ? y : cur;
            };
            if (found) {
                piv = s_piv[b];
                pivP = f.premul(piv);
#pragma unroll
                for (int k = 0; k < RPT; ++k) a[k][1] = update_entry(k, 1, stp[k] == s ? f.zero() : a[k][0], stp[k] == s, a[k][1]);
            }
            if (s + 1 < kw) {  // uniform: the search of column s + 1 starts now
                E col[RPT];
#pragma unroll
                for (int k = 0; k < RPT; ++k) col[k] = a[k][1];
                post_candidate(col, found ? p + 1 : p, b ^ 1);
            }
            if (found) {
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const bool is_piv_row = stp[k] == s;
                    const E m = is_piv_row ? f.zero() : a[k][0];
                    // groups of eight entries behind one uniform branch, selects inside: the updates of a group are
                    // independent instruction streams (a branch per entry serialises them — with one warp per scheduler
                    // nothing else hides the latency)
#pragma unroll
                    for (int g = 0; g < kPanel / 8; ++g) {
                        if (8 * g <= kPanel - 1 - s || nullmask != 0u) {  // uniform
#pragma unroll
                            for (int u = (g == 0 ? 2 : 8 * g); u < 8 * g + 8; ++u)
                                a[k][u] = update_entry(k, u, m, is_piv_row, a[k][u]);
                        }
                    }
                    if (is_piv_row) a[k][0] = f.zero();  // its own multiplier
                }
                if (tid == 0) {
                    s_pivs[s] = piv;
                    s_cs[s] = c;
                    s_pp[s] = p;
                    s_pi[s] = i;
                }
                c = f.mul(c, piv);
                swaps += i != p ? 1 : 0;
                ++p;
            } else {
                if (tid == 0) {
                    s_pivs[s] = f.zero();
                    s_cs[s] = c;
                    s_pp[s] = -1;
                    s_pi[s] = -1;
                }
                nullmask |= 1u << s;
            }
        } else if (tid == 0) {
            s_pivs[s] = f.zero();
            s_cs[s] = c;
            s_pp[s] = -1;
            s_pi[s] = -1;
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {  // rotate: the next column moves to entry 0, the finished one to the end
            const E first = a[k][0];
#pragma unroll
            for (int u = 0; u < kPanel - 1; ++u) a[k][u] = a[k][u + 1];
            a[k][kPanel - 1] = first;
        }
    }
    __syncthreads();
    // undo the scaling: 1 / c(s), the true pivot inverses c(s) / piv'(s), 1 / c(end) — 65 independent inversions
    if (tid < kPanel) {
        s_cinv[tid] = f.inv1(s_cs[tid]);
    } else if (tid < 2 * kPanel) {
        const int s = tid - kPanel;
        s_pinvt[s] = s_pp[s] >= 0 ? f.mul(s_cs[s], f.inv1(s_pivs[s])) : f.zero();
    } else if (tid == 2 * kPanel) {
        s_cendinv = f.inv1(c);
    }
    __syncthreads();
    if (rank == 0 && tid == 0) {
        E det = (E)st->det;
        for (int s = 0; s < kw; ++s)
            if (s_pp[s] >= 0) det = f.mul(det, f.mul(s_pivs[s], s_cinv[s]));
        st->det = (unsigned long long)det;
        st->p = p;
        st->n_swaps += swaps;
        meta->p0 = p0;
        meta->p1 = p;
        for (int s = 0; s < kw; ++s) {
            meta->piv_p[s] = s_pp[s];
            meta->piv_i[s] = s_pi[s];
            meta->pinv[s] = (unsigned long long)s_pinvt[s];
        }
    }
    // textbook multipliers and the finished panel columns, each row at its logical position.  A rolled loop over the columns
    // (the registers rotate as in the steps): unrolled, this block is ~10 k instructions that run once per launch at
    // instruction-fetch speed.
    E mine_pinv[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) mine_pinv[k] = stp[k] >= 0 ? s_pinvt[stp[k]] : f.one();
#pragma unroll 1
    for (int t = 0; t < kPanel; ++t) {
        if (t < kw) {  // uniform
            const bool piv_col = s_pp[t] >= 0;  // uniform
            const E scale = piv_col ? s_cinv[t] : s_cendinv;
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                if (pos[k] != INT_MAX) {
                    const size_t r = (size_t)pos[k];
                    E v = f.mul(a[k][0], scale);
                    if (stp[k] >= 0 && (!piv_col || stp[k] < t)) v = f.mul(v, mine_pinv[k]);
                    const E mult = piv_col ? v : f.zero();
                    const E entry = piv_col ? (stp[k] == t ? f.one() : f.zero()) : v;
                    M[r * kPanel + t] = mult;
                    A[r * cols + j0 + t] = to_store<T, E>(entry);
                    if (stp[k] >= 0) *cluster.map_shared_rank(&s_mp[stp[k]][t], 0) = mult;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const E first = a[k][0];
#pragma unroll
            for (int u = 0; u < kPanel - 1; ++u) a[k][u] = a[k][u + 1];
            a[k][kPanel - 1] = first;
        }
    }
    cluster.sync();
    if (rank != 0) return;
    // T: w_s = pinv_s (x_s - sum_{t<s} mp[s][t] w_t) as a matrix on x; T2 = (I - strict upper part of mp) T.  Rolled loops over
    // shared memory (unrolling the 32 x 32 recurrence is ~200 KB of straight-line code that runs at instruction-fetch speed);
    // thread (column j, quarter q): each row's dot product is split over the four warps.
    __shared__ E s_T[kPanel][kPanel + 1];
    __shared__ E s_part[4][kPanel];
    const int j = tid & 31, q = tid >> 5;
#pragma unroll 1
    for (int s = 0; s < kPanel; ++s) {
        DotAcc<F> acc;
        acc.init(f);
        for (int t = q; t < s; t += 4) acc.fma(f, s_mp[s][t], s_T[t][j]);
        s_part[q][j] = acc.done(f);
        __syncthreads();
        if (q == 0) {
            const E sum = f.add(f.add(s_part[0][j], s_part[1][j]), f.add(s_part[2][j], s_part[3][j]));
            const E v = s_pp[s] >= 0 ? f.mul(s_pinvt[s], f.sub(j == s ? f.one() : f.zero(), sum)) : f.zero();
            s_T[s][j] = v;
            TT[s * kPanel + j] = v;
        }
        __syncthreads();
    }
#pragma unroll 1
    for (int e = 0; e < kPanel / 4; ++e) {
        const int s = q * (kPanel / 4) + e;
        DotAcc<F> acc;
        acc.init(f);
#pragma unroll 4
        for (int t = s + 1; t < kPanel; ++t) acc.fma(f, s_mp[s][t], s_T[t][j]);
        TT[kPanel * kPanel + s * kPanel + j] = f.sub(s_T[s][j], acc.done(f));
    }
}

// Trailing columns after the cluster panel: exchange replay (as k_rr_panel_w), then w = T x (stored for the update kernel)
// and the finished pivot rows T2 x, x = the 32 pivot-row entries of the column.  One CTA per 32 columns; thread (column,
// quarter) computes eight rows of both products.
constexpr int kW2Cols = 32;
constexpr int kW2Threads = 128;

template <typename T, typename F, bool SMALLP = false>
__global__ void __launch_bounds__(kW2Threads)
k_rr_panel_w2(F f, T* __restrict__ A, size_t rows, size_t cols, size_t c0, size_t c_end, int kw,
              const PanelMeta* __restrict__ meta, const typename F::elem* __restrict__ TT, typename F::elem* __restrict__ W,
              size_t fs) {
    using E = typename F::elem;
    extern __shared__ uint4 smem_raw[];
    f.init_smem(smem_raw);
    (void)fs;
    (void)rows;
    __shared__ E s_T[kPanel][kPanel + 1], s_T2[kPanel][kPanel + 1], s_x[kPanel][kW2Cols + 1];
    __shared__ int s_p[kPanel], s_i[kPanel];
    __shared__ int s_pos[2 * kPanel], s_src[2 * kPanel], s_moves;
    if (meta->p1 == meta->p0) return;  // the panel found no pivot: nothing changes (uniform)
    for (int idx = threadIdx.x; idx < kPanel * kPanel; idx += blockDim.x) {
        const int s = idx / kPanel, j = idx - s * kPanel;
        s_T[s][j] = TT[idx];
        s_T2[s][j] = TT[kPanel * kPanel + idx];
    }
    if (threadIdx.x < kPanel) {
        const int s = threadIdx.x;
        s_p[s] = s < kw ? meta->piv_p[s] : -1;
        s_i[s] = s < kw ? meta->piv_i[s] : -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // net effect of the exchanges: position s_pos[k] ends up holding the row that started at s_src[k]
        int n = 0;
        for (int s = 0; s < kw; ++s) {
            const int ps = s_p[s], is = s_i[s];
            if (ps < 0 || is == ps) continue;
            int kp = -1, ki = -1;
            for (int k = 0; k < n; ++k) {
                if (s_pos[k] == ps) kp = k;
                if (s_pos[k] == is) ki = k;
            }
            if (kp < 0) {
                kp = n++;
                s_pos[kp] = ps;
                s_src[kp] = ps;
            }
            if (ki < 0) {
                ki = n++;
                s_pos[ki] = is;
                s_src[ki] = is;
            }
            const int t = s_src[kp];
            s_src[kp] = s_src[ki];
            s_src[ki] = t;
        }
        s_moves = n;
    }
    __syncthreads();
    const int col = threadIdx.x & (kW2Cols - 1), q = threadIdx.x / kW2Cols;  // q = 0..3
    const size_t c = c0 + (size_t)blockIdx.x * kW2Cols + col;
    const bool live = c < c_end;  // columns [c0, c_end)
    {
        const int n = s_moves;
        T v[2 * kPanel / 4];
#pragma unroll
        for (int e = 0; e < 2 * kPanel / 4; ++e) {
            const int k = q + 4 * e;
            if (live && k < n) v[e] = A[(size_t)s_src[k] * cols + c];
        }
        __syncthreads();  // every load of the exchange before any store
#pragma unroll
        for (int e = 0; e < 2 * kPanel / 4; ++e) {
            const int k = q + 4 * e;
            if (live && k < n && s_src[k] != s_pos[k]) A[(size_t)s_pos[k] * cols + c] = v[e];
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < kPanel / 4; ++e) {
        const int s = q * (kPanel / 4) + e;
        s_x[s][col] = (live && s_p[s] >= 0) ? to_elem<E, T>(A[(size_t)s_p[s] * cols + c]) : f.zero();
    }
    __syncthreads();
    if (!live) return;
#pragma unroll 1
    for (int e = 0; e < kPanel / 4; ++e) {
        const int s = q * (kPanel / 4) + e;
        if (s >= kw) break;
        E w = f.zero();
        if (s_p[s] >= 0) {  // uniform per warp
            E y;
            if constexpr (SMALLP) {  // p < 2^30: sixteen products per 64-bit accumulator (see k_rr_panel_update)
                unsigned long long w0 = 0ull, w1 = 0ull, y0 = 0ull, y1 = 0ull;
#pragma unroll 8
                for (int j = 0; j < kPanel / 2; ++j) {
                    const unsigned long long x0 = s_x[j][col], x1 = s_x[j + kPanel / 2][col];
                    w0 += x0 * s_T[s][j];
                    w1 += x1 * s_T[s][j + kPanel / 2];
                    y0 += x0 * s_T2[s][j];
                    y1 += x1 * s_T2[s][j + kPanel / 2];
                }
                w = f.add(f.reduce64(w0), f.reduce64(w1));
                y = f.add(f.reduce64(y0), f.reduce64(y1));
            } else {
                DotAcc<F> aw, ay;
                aw.init(f);
                ay.init(f);
#pragma unroll 8
                for (int j = 0; j < kPanel; ++j) {
                    const E x = s_x[j][col];
                    aw.fma(f, s_T[s][j], x);
                    ay.fma(f, s_T2[s][j], x);
                }
                w = aw.done(f);
                y = ay.done(f);
            }
            A[(size_t)s_p[s] * cols + c] = to_store<T, E>(y);
        }
        W[(size_t)s * cols + c] = w;
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
    PanelMeta* meta;
    void* M;  // rows x kPanel multipliers
    void* W;  // kPanel x cols scaled pivot-row entries
    void* TT; // the two kPanel x kPanel matrices of the cluster panel (T, T2)
    // second set for the pipelined cluster path: panel k + 1 is factored while the far trailing columns still read panel k's
    PanelMeta* meta2;
    void* M2;
    void* W2;
    void* TT2;
};

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

static ElimWorkspace carve(void* ws, size_t rows, size_t cols) {
    char* base = (char*)ws;
    ElimWorkspace w;
    w.state = (ElimState*)base;
    base += 64;
    w.rowbuf = base;
    base += align16(cols * 8);
    w.colbuf = base;
    base += align16(rows * 8);
    w.meta = (PanelMeta*)base;
    base += align16(sizeof(PanelMeta));
    w.M = base;
    base += align16(rows * kPanel * 8);
    w.W = base;
    base += align16(cols * kPanel * 8);
    w.TT = base;
    base += 2 * kPanel * kPanel * 8;
    w.meta2 = (PanelMeta*)base;
    base += align16(sizeof(PanelMeta));
    w.M2 = base;
    base += align16(rows * kPanel * 8);
    w.W2 = base;
    base += align16(cols * kPanel * 8);
    w.TT2 = base;
    return w;
}

// side stream + events of the pipelined cluster elimination (one set per process and device; enqueueing is serialised)
struct ElimSide {
    std::mutex mu;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_panel[2] = {nullptr, nullptr}, ev_b[2] = {nullptr, nullptr};
    cudaError_t ensure() {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (stream && dev == device) return cudaSuccess;
        if (stream) {  // another device became current: rebuild on it
            cudaStreamDestroy(stream);
            for (int i = 0; i < 2; ++i) {
                cudaEventDestroy(ev_panel[i]);
                cudaEventDestroy(ev_b[i]);
            }
            stream = nullptr;
        }
        e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
        for (int i = 0; i < 2; ++i) {
            e = cudaEventCreateWithFlags(&ev_panel[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&ev_b[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        device = dev;
        return cudaSuccess;
    }
};
static ElimSide& elim_side() {
    static ElimSide s;
    return s;
}

static size_t elim_workspace_bytes(size_t rows, size_t cols) {
    return 64 + align16(cols * 8) + align16(rows * 8) + align16(sizeof(PanelMeta)) + align16(rows * kPanel * 8) +
           align16(cols * kPanel * 8) + 2 * kPanel * kPanel * 8 + align16(sizeof(PanelMeta)) + align16(rows * kPanel * 8) +
           align16(cols * kPanel * 8) + 2 * kPanel * kPanel * 8 + 64;
}

static dim3 elim_grid(size_t rows, size_t cols, size_t c0) {
    return dim3((unsigned)((cols - c0 + kElimCols - 1) / kElimCols), (unsigned)((rows + kElimRows - 1) / kElimRows));
}

}  // namespace gfb

extern "C" {

size_t gfb_row_reduce_workspace_bytes(const gfb_field* f, size_t rows, size_t cols) {
    (void)f;
    return gfb::elim_workspace_bytes(rows, cols);
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
        ElimWorkspace w = carve(d_workspace, rows, cols);
        k_elim_init<F><<<1, 32, fs, st>>>(functor, w.state, fs);
        if (rows && cols) {
            if ((rows + kUpdRows - 1) / kUpdRows > 65535) return GFB_ERR_UNSUPPORTED;
            // the panel slab lives in shared memory when rows * 33 entries fit beside the field tables
            const size_t slab_bytes = align16(rows * kSlabStride * sizeof(SlabT<T, E, true>));
            const bool slab_in_smem = fs + slab_bytes <= (size_t)200 * 1024;
            if (slab_in_smem)
                GFB_CUDA_TRY(cudaFuncSetAttribute(k_rr_panel<T, F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)(fs + slab_bytes)));
            // fields with 32-bit elements, up to 4096 rows: the panel runs on a thread-block cluster (k_rr_panel_cluster)
            // and the trailing pivot rows come from its T / T2 matrices (k_rr_panel_w2).  GFB_ELIM_CLUSTER=0 keeps the
            // single-CTA panel (the tests run both and compare).
            bool use_cluster = false;
            if constexpr (sizeof(E) == 4) {
                const char* env = getenv("GFB_ELIM_CLUSTER");  // read per call: the tests switch it inside one process
                const bool cluster_on = !(env && env[0] == '0');
                use_cluster = cluster_on && rows <= (size_t)4 * kClRows && fs <= (size_t)160 * 1024;
            }
            if constexpr (sizeof(E) == 4) {
                if (use_cluster) {
                    // Pipelined: after panel k only the columns of panel k + 1 are updated on the caller's stream (part a);
                    // the far trailing columns (part b) run on a side stream WHILE panel k + 1 is factored on its 8 SMs.
                    // meta / M / W / TT alternate between two sets so that part b of panel k and panel k + 1 never share one.
                    ElimSide& side = elim_side();
                    std::lock_guard<std::mutex> lock(side.mu);
                    GFB_CUDA_TRY(side.ensure());
                    bool small_p = false;
                    if constexpr (std::is_same_v<F, PrimeBarrett>) small_p = functor.p < (1u << 30);
                    auto trailing = [&](cudaStream_t s2, size_t cb, size_t ce, int kw, int q) {
                        if (cb >= ce) return;
                        PanelMeta* meta = q ? w.meta2 : w.meta;
                        const E* Mq = (const E*)(q ? w.M2 : w.M);
                        E* Wq = (E*)(q ? w.W2 : w.W);
                        const E* TTq = (const E*)(q ? w.TT2 : w.TT);
                        const unsigned gw = (unsigned)((ce - cb + kW2Cols - 1) / kW2Cols);
                        dim3 gu((unsigned)((ce - cb + kUpdCols - 1) / kUpdCols), (unsigned)((rows + kUpdRows - 1) / kUpdRows));
                        if constexpr (std::is_same_v<F, PrimeBarrett>) {
                            if (small_p) {
                                k_rr_panel_w2<T, F, true><<<gw, kW2Threads, fs, s2>>>(functor, (T*)A, rows, cols, cb, ce, kw, meta, TTq, Wq, fs);
                                k_rr_panel_update<T, F, true><<<gu, kUpdCols, fs, s2>>>(functor, (T*)A, rows, cols, cb, ce, kw, meta, Mq, Wq, fs);
                                return;
                            }
                        }
                        k_rr_panel_w2<T, F><<<gw, kW2Threads, fs, s2>>>(functor, (T*)A, rows, cols, cb, ce, kw, meta, TTq, Wq, fs);
                        k_rr_panel_update<T, F><<<gu, kUpdCols, fs, s2>>>(functor, (T*)A, rows, cols, cb, ce, kw, meta, Mq, Wq, fs);
                    };
                    if (fs > 48 * 1024) {
                        GFB_CUDA_TRY(cudaFuncSetAttribute(k_rr_panel_cluster<T, F, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
                        GFB_CUDA_TRY(cudaFuncSetAttribute(k_rr_panel_cluster<T, F, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
                    }
                    bool side_used = false;
                    size_t k = 0;
                    for (size_t j0 = 0; j0 < ncols; j0 += kPanel, ++k) {
                        const int kw = (int)(ncols - j0 < (size_t)kPanel ? ncols - j0 : (size_t)kPanel);
                        const int q = (int)(k & 1);
                        PanelMeta* meta = q ? w.meta2 : w.meta;
                        E* Mq = (E*)(q ? w.M2 : w.M);
                        E* TTq = (E*)(q ? w.TT2 : w.TT);
                        bool odd_p = false;
                        if constexpr (std::is_same_v<F, PrimeBarrett>) odd_p = (functor.p & 1u) != 0;
                        if constexpr (std::is_same_v<F, PrimeBarrett>) {
                            if (odd_p && rows <= (size_t)kClRows)
                                k_rr_panel_cluster<T, F, 1, true><<<kClCtas, kClThreads, fs, st>>>(functor, (T*)A, rows, cols, j0, kw, w.state, meta, Mq, TTq, fs);
                            else if (odd_p)
                                k_rr_panel_cluster<T, F, 4, true><<<kClCtas, kClThreads, fs, st>>>(functor, (T*)A, rows, cols, j0, kw, w.state, meta, Mq, TTq, fs);
                        }
                        if (!odd_p) {
                            if (rows <= (size_t)kClRows)
                                k_rr_panel_cluster<T, F, 1><<<kClCtas, kClThreads, fs, st>>>(functor, (T*)A, rows, cols, j0, kw, w.state, meta, Mq, TTq, fs);
                            else
                                k_rr_panel_cluster<T, F, 4><<<kClCtas, kClThreads, fs, st>>>(functor, (T*)A, rows, cols, j0, kw, w.state, meta, Mq, TTq, fs);
                        }
                        const size_t c0 = j0 + kw;  // trailing columns [c0, cols)
                        const bool more = c0 < ncols;  // another panel follows: its columns are [c0, c0 + 32)
                        const size_t ca = more ? (c0 + kPanel < cols ? c0 + kPanel : cols) : cols;
                        // part b of the previous panel also covers [c0, cols): it must be finished before part a touches them
                        if (side_used) GFB_CUDA_TRY(cudaStreamWaitEvent(st, side.ev_b[q ^ 1], 0));
                        if (more && ca < cols) {
                            GFB_CUDA_TRY(cudaEventRecord(side.ev_panel[q], st));
                            GFB_CUDA_TRY(cudaStreamWaitEvent(side.stream, side.ev_panel[q], 0));
                            trailing(side.stream, ca, cols, kw, q);
                            GFB_CUDA_TRY(cudaEventRecord(side.ev_b[q], side.stream));
                            side_used = true;
                            trailing(st, c0, ca, kw, q);
                        } else {
                            side_used = false;
                            trailing(st, c0, cols, kw, q);
                        }
                    }
                    if (side_used) GFB_CUDA_TRY(cudaStreamWaitEvent(st, side.ev_b[(int)((k - 1) & 1)], 0));
                }
            }
            for (size_t j0 = 0; j0 < ncols && !use_cluster; j0 += kPanel) {
                const int kw = (int)(ncols - j0 < (size_t)kPanel ? ncols - j0 : (size_t)kPanel);
                const size_t c0 = j0 + kw;  // trailing columns
                if (slab_in_smem)
                    k_rr_panel<T, F, true><<<1, kPivotThreads, fs + slab_bytes, st>>>(functor, (T*)A, rows, cols, j0, kw,
                                                                                       w.state, w.meta, (E*)w.M, fs);
                else
                    k_rr_panel<T, F, false><<<1, kPivotThreads, fs, st>>>(functor, (T*)A, rows, cols, j0, kw, w.state,
                                                                           w.meta, (E*)w.M, fs);
                if (c0 < cols) {
                    const unsigned gw = (unsigned)((cols - c0 + kPanelWThreads - 1) / kPanelWThreads);
                    k_rr_panel_w<T, F><<<gw, kPanelWThreads, fs, st>>>(functor, (T*)A, rows, cols, c0, kw, w.meta,
                                                                       (const E*)w.M, (E*)w.W, fs);
                    dim3 gu((unsigned)((cols - c0 + kUpdCols - 1) / kUpdCols), (unsigned)((rows + kUpdRows - 1) / kUpdRows));
                    k_rr_panel_update<T, F><<<gu, kUpdCols, fs, st>>>(functor, (T*)A, rows, cols, c0, cols, kw, w.meta,
                                                                      (const E*)w.M, (const E*)w.W, fs);
                }
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
    if (!pivoting && rows - 1 > cols) return GFB_ERR_BAD_ARG;  // the diagonal walk of lu_decompose needs rows - 1 <= cols
    if (rows > (size_t)INT_MAX / 2 || cols > (size_t)INT_MAX / 2) return GFB_ERR_UNSUPPORTED;
    if ((rows + kElimRows - 1) / kElimRows > 65535) return GFB_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch_field(f, eb, [&](auto functor, auto tag) -> gfb_status {
        using F = decltype(functor);
        using T = typename decltype(tag)::type;
        using E = typename F::elem;
        const size_t fs = field_smem_bytes(functor);
        ElimWorkspace w = carve(d_workspace, rows, cols);
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
