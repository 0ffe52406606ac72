// gfb_ewise.cu — element-wise GF(p) / GF(2^m) ufunc kernels (add, sub, mul, div, neg, inv,
// square, power) for sm_100a.
//
// These replace the reference's numba.vectorize DUFuncs (one scalar `calculate`/`lookup`
// per element, /root/reference/src/galois/_domains/_ufunc.py:97-144) with HBM-streaming
// kernels: native-width 128-bit loads (the reference round-trips through int64 twice,
// _ufunc.py:302-330), arithmetic in registers, 128-bit evict-first stores.  They are
// integer, memory-bound kernels: no tensor cores, no shared memory except the EXP/LOG
// tables of small binary fields.
//
// Algorithmic bytes per element: (2 reads + 1 write) * elem_bytes for binary ops,
// (1 read + 1 write) * elem_bytes for unary ops.
//
// Build structure: the kernel templates live in gfb_ewise_impl.cuh and are instantiated once per field functor in
// gfb_ewise_{m31,barrett,lut,calc,b128}.cu (parallel compilation); this file keeps the C-ABI entry points and the
// dispatch on the field implementation.
#include "gfb_common.cuh"
#include "gfb_ewise_api.cuh"

namespace gfb {

// does an element of this field fit the storage width?
static bool storage_ok(const gfb_field* f, int eb) {
    if (f->impl == IMPL_BIN_128) return eb == 16;
    if (eb != 1 && eb != 2 && eb != 4 && eb != 8) return false;
    if (eb == 8) return true;
    const int bits = 8 * eb;
    if (f->kind == GFB_PRIME) return (f->order_lo - 1) < (1ull << bits);
    return f->degree <= bits;
}

static gfb_status binary_dispatch(gfb_field* f, int op, int mode, int eb, const void* a, const void* b, uint64_t s_lo,
                                  uint64_t s_hi, void* out, size_t n, void* stream) {
    if (!f || !a || !out || (mode == MODE_AA && !b)) return n == 0 && f ? GFB_OK : GFB_ERR_BAD_ARG;
    if (op < GFB_ADD || op > GFB_DIV) return GFB_ERR_BAD_ARG;
    if (!storage_ok(f, eb)) return GFB_ERR_BAD_ARG;
    if (n == 0) return GFB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (f->impl) {
        case IMPL_PRIME_M31:
            if (s_lo >= f->order_lo) return GFB_ERR_BAD_ARG;
            return ewise_binary_m31(f->prime_m31(), op, mode, eb, a, b, s_lo, s_hi, out, n, f->d_err, st);
        case IMPL_PRIME_BARRETT:
            if (mode != MODE_AA && s_lo >= f->order_lo) return GFB_ERR_BAD_ARG;
            return ewise_binary_barrett(f->prime_barrett(), op, mode, eb, a, b, s_lo, s_hi, out, n, f->d_err, st);
        case IMPL_BIN_LUT:
            if (mode != MODE_AA && s_lo >= f->order_lo) return GFB_ERR_BAD_ARG;
            if (tab8_applies(f, op, eb, a, mode == MODE_AA ? b : a, out))
                return ewise_binary_tab8(f, op, mode, a, b, s_lo, out, n, st);
            return ewise_binary_lut(f->bin_lut(), op, mode, eb, a, b, s_lo, s_hi, out, n, f->d_err, st);
        case IMPL_BIN_CALC:
            if (mode != MODE_AA && f->degree < 64 && s_lo >= f->order_lo) return GFB_ERR_BAD_ARG;
            return ewise_binary_calc(f->bin_calc(), op, mode, eb, a, b, s_lo, s_hi, out, n, f->d_err, st);
        case IMPL_BIN_128:
            return ewise_binary_b128(f->bin_128(), op, mode, eb, a, b, s_lo, s_hi, out, n, f->d_err, st);
    }
    return GFB_ERR_BAD_ARG;
}

static gfb_status power_dispatch(gfb_field* f, int eb, const void* a, const int64_t* e, int64_t es, bool scalar_e,
                                 void* out, size_t n, void* stream) {
    if (!f) return GFB_ERR_BAD_ARG;
    if (n == 0) return GFB_OK;
    if (!a || !out || (!scalar_e && !e)) return GFB_ERR_BAD_ARG;
    if (!storage_ok(f, eb)) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (f->impl) {
        case IMPL_PRIME_M31: return ewise_power_m31(f->prime_m31(), eb, a, e, es, scalar_e, out, n, f->d_err, st);
        case IMPL_PRIME_BARRETT: return ewise_power_barrett(f->prime_barrett(), eb, a, e, es, scalar_e, out, n, f->d_err, st);
        case IMPL_BIN_LUT: return ewise_power_lut(f->bin_lut(), eb, a, e, es, scalar_e, out, n, f->d_err, st);
        case IMPL_BIN_CALC: return ewise_power_calc(f->bin_calc(), eb, a, e, es, scalar_e, out, n, f->d_err, st);
        case IMPL_BIN_128:
            return ewise_power_b128(f->bin_128(), eb, a, e, es, scalar_e, out, n, f->d_err, st);
    }
    return GFB_ERR_BAD_ARG;
}

}  // namespace gfb

extern "C" {

gfb_status gfb_binary(gfb_field* f, int op, int elem_bytes, const void* a, const void* b, void* out, size_t n,
                      void* stream) {
    return gfb::binary_dispatch(f, op, gfb::MODE_AA, elem_bytes, a, b, 0, 0, out, n, stream);
}

gfb_status gfb_binary_scalar(gfb_field* f, int op, int elem_bytes, const void* a, uint64_t s_lo, uint64_t s_hi,
                             int scalar_on_left, void* out, size_t n, void* stream) {
    return gfb::binary_dispatch(f, op, scalar_on_left ? gfb::MODE_SA : gfb::MODE_AS, elem_bytes, a, nullptr, s_lo, s_hi,
                                out, n, stream);
}

gfb_status gfb_unary(gfb_field* f, int op, int elem_bytes, const void* a, void* out, size_t n, void* stream) {
    using namespace gfb;
    if (!f) return GFB_ERR_BAD_ARG;
    if (n == 0) return GFB_OK;
    if (!a || !out) return GFB_ERR_BAD_ARG;
    if (op < GFB_NEG || op > GFB_SQUARE) return GFB_ERR_BAD_ARG;
    if (!storage_ok(f, elem_bytes)) return GFB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (f->impl) {
        case IMPL_PRIME_M31: return ewise_unary_m31(f->prime_m31(), op, elem_bytes, a, out, n, f->d_err, st);
        case IMPL_PRIME_BARRETT: return ewise_unary_barrett(f->prime_barrett(), op, elem_bytes, a, out, n, f->d_err, st);
        case IMPL_BIN_LUT: return ewise_unary_lut(f->bin_lut(), op, elem_bytes, a, out, n, f->d_err, st);
        case IMPL_BIN_CALC: return ewise_unary_calc(f->bin_calc(), op, elem_bytes, a, out, n, f->d_err, st);
        case IMPL_BIN_128: return ewise_unary_b128(f->bin_128(), op, elem_bytes, a, out, n, f->d_err, st);
    }
    return GFB_ERR_BAD_ARG;
}

gfb_status gfb_power(gfb_field* f, int elem_bytes, const void* a, const int64_t* e, void* out, size_t n, void* stream) {
    return gfb::power_dispatch(f, elem_bytes, a, e, 0, false, out, n, stream);
}

gfb_status gfb_power_scalar(gfb_field* f, int elem_bytes, const void* a, int64_t e, void* out, size_t n, void* stream) {
    return gfb::power_dispatch(f, elem_bytes, a, nullptr, e, true, out, n, stream);
}

}  // extern "C"
