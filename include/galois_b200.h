/*
 * galois_b200.h — C ABI of the B200-native finite-field backend.
 *
 * This is the drop-in boundary: every entry point below replaces one dispatcher
 * of mhostetter/galois's L2 "UFunc"/"Function" layer (the reference has no FFI of
 * its own — its "plugin API" is two Python dispatcher classes).
 * The reference interface each function replaces is cited as file:line under
 * /root/reference/src/galois/.
 *
 * Conventions
 *   - plain C: pointers, sizes, integers; no torch / C++ types cross this boundary.
 *   - `*_dev` style functions (no suffix) take DEVICE pointers and a `stream`
 *     (a cudaStream_t cast to void*; NULL = legacy default stream); they are
 *     asynchronous and never synchronise.
 *   - `*_host` functions take HOST pointers, run a chunked H2D → kernel → D2H
 *     pipeline on internal streams and return after the result is in `out`.
 *   - element storage is native width: elem_bytes ∈ {1,2,4,8} (the reference's
 *     dtypes uint8/uint16/uint32/int8/int16/int32/int64, _domains/_meta.py:19;
 *     values are always non-negative so signedness does not matter), or 16 for
 *     GF(2^m), 64 <= m <= 128 (two little-endian uint64 limbs per element; the
 *     reference holds these as Python ints in dtype=object arrays).
 *   - every function returns a gfb_status; nothing throws across the boundary.
 *   - "zero division" (reference: ZeroDivisionError raised from inside the Numba
 *     loop, _domains/_calculate.py:403-404, _lookup.py:194-195,225-226,261-262)
 *     is a sticky per-context flag set by the kernel and read with
 *     gfb_take_error(); `*_host` functions return GFB_ERR_ZERO_DIVISION directly.
 */
#ifndef GALOIS_B200_H
#define GALOIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum gfb_status {
    GFB_OK = 0,
    GFB_ERR_ZERO_DIVISION = 1, /* inverse of 0 / 0**negative requested            */
    GFB_ERR_BAD_ARG = 2,       /* unsupported field / dtype / size / null pointer */
    GFB_ERR_CUDA = 3,          /* a CUDA runtime call failed; see gfb_last_error  */
    GFB_ERR_UNSUPPORTED = 4    /* valid in the reference, not built here (yet)    */
} gfb_status;

typedef enum gfb_field_kind {
    GFB_PRIME = 1,  /* GF(p), p prime, p < 2^32            (_fields/_ufunc.py:20-33)  */
    GFB_BINARY = 2  /* GF(2^m), 1 <= m <= 128              (_fields/_ufunc.py:56-74)  */
} gfb_field_kind;

/* Binary ufuncs (UFunc dispatchers, _domains/_ufunc.py:339-441). */
typedef enum gfb_binop {
    GFB_ADD = 0, /* add_modular _calculate.py:133-147 / np.bitwise_xor _fields/_ufunc.py:59 */
    GFB_SUB = 1, /* subtract_modular _calculate.py:235-251 / xor                            */
    GFB_MUL = 2, /* multiply_modular :327-340, multiply_binary :288-324, lookup _lookup.py:143-168 */
    GFB_DIV = 3  /* divide _calculate.py:492-513, _lookup.py:201-233; one fused pass        */
} gfb_binop;

/* Unary ufuncs. */
typedef enum gfb_unop {
    GFB_NEG = 0,   /* negative_modular _calculate.py:184-199 / np.positive for GF(2^m) */
    GFB_INV = 1,   /* reciprocal_modular_egcd :386-417, reciprocal_itoh_tsujii :447-489 */
    GFB_SQUARE = 2 /* square_ufunc _domains/_ufunc.py:493-503 (x*x)                    */
} gfb_unop;

typedef struct gfb_field gfb_field;       /* opaque: constants + device tables           */
typedef struct gfb_ntt_plan gfb_ntt_plan; /* opaque: twiddle tables for one (field,N,ω)  */
typedef struct gfb_rs_code gfb_rs_code;   /* opaque: RS(n,k) generator/roots on device   */

/* ---------------------------------------------------------------- library ---- */
/* Number of CUDA devices visible, or 0 (never fails). */
int gfb_device_count(void);
/* Human-readable text of the last failing CUDA call on this thread ("" if none). */
const char* gfb_last_error(void);
/* Version string "galois_b200 x.y (sm_100a)". */
const char* gfb_version(void);
/* Read-and-clear the sticky error flag of `f` after synchronising `stream`.
 * Returns GFB_OK or GFB_ERR_ZERO_DIVISION. */
gfb_status gfb_take_error(gfb_field* f, void* stream);

/* Pinned host memory for the `*_host` pipelines (cudaHostAlloc / cudaFreeHost). */
void* gfb_host_alloc(size_t bytes);
void gfb_host_free(void* p);

/* ------------------------------------------------------------------ fields ---- */
/* Replaces FieldArrayMeta/`Array.compile` (_domains/_array.py:322-362) + table build
 * `_build_lookup_tables` (_domains/_lookup.py:319-371).
 *   kind=GFB_PRIME : order_lo = p (p < 2^32), degree = 1, poly_* ignored.
 *   kind=GFB_BINARY: degree = m; poly_lo/poly_hi = low 128 bits of the irreducible
 *                    polynomial's integer representation INCLUDING x^m when m < 128
 *                    (for m = 128 the x^128 term is implicit); primitive_element is
 *                    the generator used for the EXP/LOG tables (m <= 16 only).
 * The field is bound to the CUDA device current at creation. */
gfb_status gfb_field_create(gfb_field** out, int kind, uint64_t order_lo, int degree,
                            uint64_t poly_lo, uint64_t poly_hi, uint64_t primitive_element);
void gfb_field_destroy(gfb_field* f);
/* 1 if the EXP/LOG tables are in use (m <= 12: staged in shared memory). */
int gfb_field_uses_tables(const gfb_field* f);

/* --------------------------------------------------------- element-wise ufuncs */
/* out[i] = a[i] (op) b[i], i < n.   UFunc.__call__ on two arrays, _ufunc.py:339-441. */
gfb_status gfb_binary(gfb_field* f, int op, int elem_bytes, const void* a, const void* b,
                      void* out, size_t n, void* stream);
/* out[i] = a[i] (op) s   (scalar_on_left=0)   or   s (op) a[i]   (scalar_on_left=1).
 * s is a field element given as (lo, hi) limbs (hi = 0 unless elem_bytes == 16).
 * NumPy broadcasting of a 0-d operand, _ufunc.py:391-401. */
gfb_status gfb_binary_scalar(gfb_field* f, int op, int elem_bytes, const void* a, uint64_t s_lo,
                             uint64_t s_hi, int scalar_on_left, void* out, size_t n, void* stream);
/* out[i] = op(a[i]). */
gfb_status gfb_unary(gfb_field* f, int op, int elem_bytes, const void* a, void* out, size_t n,
                     void* stream);
/* out[i] = a[i] ** e[i], e int64 (negative allowed).  power_square_and_multiply
 * _calculate.py:558-592, power_ufunc.lookup _lookup.py:236-270. */
gfb_status gfb_power(gfb_field* f, int elem_bytes, const void* a, const int64_t* e, void* out,
                     size_t n, void* stream);
/* out[i] = a[i] ** e (one exponent for the whole array). */
gfb_status gfb_power_scalar(gfb_field* f, int elem_bytes, const void* a, int64_t e, void* out,
                            size_t n, void* stream);

/* Host-buffer variants: same semantics, pointers are host memory (pinned is faster).
 * Chunked, double-buffered H2D/compute/D2H on internal streams. */
gfb_status gfb_binary_host(gfb_field* f, int op, int elem_bytes, const void* a, const void* b,
                           void* out, size_t n);
gfb_status gfb_binary_scalar_host(gfb_field* f, int op, int elem_bytes, const void* a,
                                  uint64_t s_lo, uint64_t s_hi, int scalar_on_left, void* out,
                                  size_t n);
gfb_status gfb_unary_host(gfb_field* f, int op, int elem_bytes, const void* a, void* out, size_t n);
gfb_status gfb_power_host(gfb_field* f, int elem_bytes, const void* a, const int64_t* e, void* out,
                          size_t n);
gfb_status gfb_power_scalar_host(gfb_field* f, int elem_bytes, const void* a, int64_t e, void* out,
                                 size_t n);

/* ------------------------------------------------------------ ufunc methods ---- */
/* np.<ufunc>.reduce / np.<ufunc>.accumulate on a field array: the reference forwards ufunc methods to the compiled
 * DUFunc (UFunc dispatchers, _domains/_ufunc.py:349,403; tests/fields/test_numpy_ufuncs.py) — a serial fold with the
 * field's add / multiply.  `a` is viewed as [outer][len][inner] (C order: any axis of an N-D array); op is GFB_ADD or
 * GFB_MUL (subtract / divide folds are a0 - (a1 + ... ) and a0 / (a1 * ...): the caller composes them).
 *   scan = 0: out[outer][inner]      = a[o][0][i] (op) a[o][1][i] (op) ...      (reduce)
 *   scan = 1: out[outer][len][inner] = running fold along len                   (accumulate)
 * The device variant needs gfb_fold_workspace_bytes() of scratch (0 for most shapes).  64 < degree fields: UNSUPPORTED. */
size_t gfb_fold_workspace_bytes(int elem_bytes, size_t outer, size_t len, size_t inner);
gfb_status gfb_fold(gfb_field* f, int op, int scan, int elem_bytes, const void* a, void* out, size_t outer, size_t len,
                    size_t inner, void* workspace, void* stream);
gfb_status gfb_fold_host(gfb_field* f, int op, int scan, int elem_bytes, const void* a, void* out, size_t outer, size_t len,
                         size_t inner);
/* np.<ufunc>.reduceat: out[s] = fold of a[start[s] .. end[s]) for n_seg non-empty segments of the 1-D array a
 * (start / end: int64, device memory for the device variant). */
gfb_status gfb_fold_segments(gfb_field* f, int op, int elem_bytes, const void* a, const int64_t* start, const int64_t* end,
                             void* out, size_t n_seg, void* stream);
gfb_status gfb_fold_segments_host(gfb_field* f, int op, int elem_bytes, const void* a, size_t n, const int64_t* start,
                                  const int64_t* end, void* out, size_t n_seg);

/* ----------------------------------------------------------------- NTT / INTT */
/* Replaces fft_jit / ifft_jit (_domains/_function.py:170-392) behind galois.ntt/intt
 * (_ntt.py:16-278) and np.fft.fft/ifft on field arrays, and ADDS the batch dimension
 * the reference lacks (it is 1-D only, _function.py:180-181).
 *   X[b][k] = scale * sum_j x[b][j] * omega^(j*k)  in the field, natural order in and out.
 * `n` is any size with 1 <= n <= 2^24 and n | order-1:
 *   - powers of two over GF(p) take the radix-2 / four-step kernels (_function.py:337-352),
 *   - every other size, and GF(2^m) with m <= 63, take the reference's general mixed-radix
 *     stages (_function.py:354-376; prime factors of n with multiplicity, consumed from the end).
 * `omega` is a primitive n-th root of unity chosen by the caller exactly as the reference does
 * (FieldArray.primitive_root_of_unity, _fields/_array.py:1127-1187; its inverse for the inverse
 * transform, _function.py:194-195) — GFB_ERR_BAD_ARG if it does not have order n; `scale` is 1 or
 * (n mod p)^-1 (_function.py:209-210). */
gfb_status gfb_ntt_plan_create(gfb_ntt_plan** out, gfb_field* f, size_t n, uint64_t omega,
                               uint64_t scale);
void gfb_ntt_plan_destroy(gfb_ntt_plan* plan);
/* Bytes of device workspace gfb_ntt_exec needs for `batch` transforms (0 if none). */
size_t gfb_ntt_workspace_bytes(const gfb_ntt_plan* plan, int elem_bytes, size_t batch);
/* x, y: [batch][n] contiguous device arrays of elem_bytes ∈ {4, 8}; y may alias x, except for a prime
 * non-power-of-two n (a single out-of-place stage: GFB_ERR_BAD_ARG). */
gfb_status gfb_ntt_exec(gfb_ntt_plan* plan, int elem_bytes, const void* x, void* y, size_t batch,
                        void* workspace, void* stream);
gfb_status gfb_ntt_exec_host(gfb_ntt_plan* plan, int elem_bytes, const void* x, void* y,
                             size_t batch);

/* -------------------------------------------------- convolution and matmul ---- */
/* Full linear convolution, out has na + nb - 1 elements.  convolve_jit
 * (_domains/_function.py:111-167). */
gfb_status gfb_convolve(gfb_field* f, int elem_bytes, const void* a, size_t na, const void* b,
                        size_t nb, void* out, void* stream);
gfb_status gfb_convolve_host(gfb_field* f, int elem_bytes, const void* a, size_t na, const void* b,
                             size_t nb, void* out);
/* y[i] = coeffs[0] x[i]^(nc-1) + ... + coeffs[nc-1]: one polynomial (degree-descending coefficients) evaluated at n
 * points by Horner's rule.  evaluate_elementwise_jit (_polys/_dense.py:404-440) behind Poly.__call__
 * (SURVEY.md §8f rank 2). */
gfb_status gfb_poly_evaluate(gfb_field* f, int elem_bytes, const void* coeffs, size_t nc, const void* x,
                             void* y, size_t n, void* stream);
gfb_status gfb_poly_evaluate_host(gfb_field* f, int elem_bytes, const void* coeffs, size_t nc, const void* x,
                                  void* y, size_t n);
/* All roots of f(x) = sum_j coeffs[j] x^degrees[j] in the field, in increasing order — roots_jit
 * (_polys/_dense.py:443-513: Chien search + argsort) behind Poly.roots (_polys/_poly.py:711-790).
 * `degrees` (HOST array, strictly descending, >= 0) and `coeffs` (non-zero coefficients, device / host per variant)
 * are Poly.nonzero_degrees / Poly.nonzero_coeffs.  At most max_roots roots are written; *n_roots receives how many
 * the polynomial has (<= its degree for a non-zero polynomial).  Fields with more than 2^33 elements return
 * GFB_ERR_UNSUPPORTED (the search is exhaustive, as the reference's is).  The device variant needs
 * gfb_poly_roots_workspace_bytes() of device scratch and writes the count to device memory. */
size_t gfb_poly_roots_workspace_bytes(const gfb_field* f, size_t n_terms);
gfb_status gfb_poly_roots(gfb_field* f, int elem_bytes, const int64_t* degrees, const void* coeffs,
                          size_t n_terms, void* roots, size_t max_roots, unsigned long long* n_roots,
                          void* workspace, void* stream);
gfb_status gfb_poly_roots_host(gfb_field* f, int elem_bytes, const int64_t* degrees, const void* coeffs,
                               size_t n_terms, void* roots, size_t max_roots, size_t* n_roots);
/* C[batch][M][N] = A[batch][M][K] · B[batch or 1][K][N] over the field.  matmul_jit
 * (_domains/_linalg.py:180-308).  b_batched = 0 broadcasts one B to every batch. */
gfb_status gfb_matmul(gfb_field* f, int elem_bytes, const void* A, const void* B, void* C,
                      size_t batch, size_t M, size_t K, size_t N, int b_batched, void* stream);
gfb_status gfb_matmul_host(gfb_field* f, int elem_bytes, const void* A, const void* B, void* C,
                           size_t batch, size_t M, size_t K, size_t N, int b_batched);

/* ------------------------------------------------- Gaussian elimination ---- */
/* Reduced row echelon form of the row-major rows x cols matrix A, in place, with pivots searched in the first
 * `ncols` columns only — row_reduce_jit.__call__ (_domains/_linalg.py:321-353), including its pivoting rule (first
 * non-zero entry at or below the pivot row), which the [A | I] callers depend on: inv_jit (:492-522), solve_jit
 * (:525-546), matrix_rank_jit (:479-489), FieldArray.row_space / column_space / left_null_space / null_space
 * (_fields/_array.py:1541-1755).  SURVEY.md §8f rank 4.
 *   info (device, 2 x int64): [0] = number of pivots p, [1] = number of row exchanges.
 *   det  (device, elem_bytes, may be NULL): (-1)^exchanges x product of the pivots — det(A) when A is square,
 *        ncols == cols and p == rows (the value det_jit :446-471 computes through PLU; 0 otherwise is the caller's job).
 * The device variant needs gfb_row_reduce_workspace_bytes() of scratch.  Fields wider than 64 bits: UNSUPPORTED.
 * Fields whose elements fit 32 bits, up to 4096 rows: the 32-column panel runs on a thread-block cluster and the far
 * trailing columns of a panel are updated on an internal side stream while the next panel is factored; the side stream
 * is joined back into `stream` before the call returns, so the caller sees ordinary stream semantics
 * (GFB_ELIM_CLUSTER=0 keeps the single-CTA panel). */
size_t gfb_row_reduce_workspace_bytes(const gfb_field* f, size_t rows, size_t cols);
gfb_status gfb_row_reduce(gfb_field* f, int elem_bytes, void* A, size_t rows, size_t cols, size_t ncols,
                          int64_t* info, void* det, void* workspace, void* stream);
gfb_status gfb_row_reduce_host(gfb_field* f, int elem_bytes, const void* A, void* A_rre, size_t rows,
                               size_t cols, size_t ncols, int64_t* n_pivots, void* det);
/* A = P L U.  pivoting != 0: plu_decompose_jit.__call__ (_domains/_linalg.py:390-428) — rows are exchanged only when
 * the diagonal entry is zero, with the first non-zero entry below it; pivoting == 0: lu_decompose_jit.__call__
 * (:356-387), which fails (info[1] = 1; the reference raises ValueError) when an exchange would be needed and requires
 * rows - 1 <= cols.  A (rows x cols) becomes U in place, L is rows x rows, perm[i] (int64) is the row of A that ended
 * up as row i (the reference's row-swapped identity has a 1 at [i, perm[i]]; it returns the transpose).
 *   info (device, 2 x int64): [0] = number of row exchanges, [1] = LU-without-pivoting failed.
 * Same workspace as gfb_row_reduce. */
gfb_status gfb_plu_decompose(gfb_field* f, int elem_bytes, void* A, void* L, int64_t* perm, size_t rows,
                             size_t cols, int pivoting, int64_t* info, void* workspace, void* stream);
gfb_status gfb_plu_decompose_host(gfb_field* f, int elem_bytes, const void* A, void* U, void* L,
                                  int64_t* perm, size_t rows, size_t cols, int pivoting,
                                  int64_t* n_swaps, int* failed);

/* ------------------------------------------------------------ Reed-Solomon ---- */
/* Replaces ReedSolomon.__init__ (_codes/_reed_solomon.py:111-218) +
 * _poly_to_generator_matrix (_codes/_cyclic.py:198-226, systematic) for codes over
 * GF(2^m), m <= 8: generator roots alpha^c .. alpha^(c+d-2), d = n-k+1. */
gfb_status gfb_rs_create(gfb_rs_code** out, gfb_field* f, int n, int k, int c, uint64_t alpha);
/* Binary BCH(n, k) with design distance d over GF(2), syndromes in extension_field = GF(2^m), m <= 8: replaces
 * BCH.__init__ (_codes/_bch.py:184-240).  `generator_poly` holds the n-k+1 coefficients (0/1, degree-descending)
 * of g(x) = lcm of the minimal polynomials of alpha^c .. alpha^(c+d-2) (_bch.py:1178-1197), computed by the
 * caller; the returned handle is used with the gfb_rs_* encode / decode / get_* entry points (the reference
 * shares one decoder between both code families, _reed_solomon.py:1105-1113).  Symbols are bytes 0/1. */
gfb_status gfb_bch_create(gfb_rs_code** out, gfb_field* extension_field, int n, int k, int d, int c,
                          uint64_t alpha, const uint8_t* generator_poly);
void gfb_rs_destroy(gfb_rs_code* rs);
/* Copies of the code's constants (host memory, caller-allocated):
 *   generator_poly: n-k+1 coefficients, degree-descending;  parity: k x (n-k) matrix P of G=[I|P]. */
gfb_status gfb_rs_get_generator_poly(const gfb_rs_code* rs, uint8_t* generator_poly);
gfb_status gfb_rs_get_parity_matrix(const gfb_rs_code* rs, uint8_t* parity);
/* Systematic encode: msg [N][ks] → codeword [N][ks + (n-k)], ks <= k (shortened codes pad
 * leading zeros).  _LinearCode._encode_message (_codes/_linear.py:270-284). */
gfb_status gfb_rs_encode(gfb_rs_code* rs, const uint8_t* msg, int ks, uint8_t* codeword, size_t N,
                         void* stream);
/* Errors-and-erasures decode: received [N][ns] (+ optional erasure mask [N][ns], bytes 0/1, may
 * be NULL) → corrected codewords [N][ns] and n_errors [N] (int64; -1 = uncorrectable, row left
 * equal to the received row).  bch_decode_jit.implementation (_codes/_bch.py:1337-1578).
 * Codes with at most 32 parity symbols and N >= 32 run the bit-sliced kernel (32 codewords per warp); the buffers may
 * start at any byte offset.  GFB_RS_DECODER=log|fast|bs selects a kernel explicitly (same results). */
gfb_status gfb_rs_decode(gfb_rs_code* rs, const uint8_t* received, const uint8_t* erasures, int ns,
                         uint8_t* decoded, int64_t* n_errors, size_t N, void* stream);
gfb_status gfb_rs_encode_host(gfb_rs_code* rs, const uint8_t* msg, int ks, uint8_t* codeword,
                              size_t N);
gfb_status gfb_rs_decode_host(gfb_rs_code* rs, const uint8_t* received, const uint8_t* erasures,
                              int ns, uint8_t* decoded, int64_t* n_errors, size_t N);

#ifdef __cplusplus
}
#endif
#endif /* GALOIS_B200_H */
