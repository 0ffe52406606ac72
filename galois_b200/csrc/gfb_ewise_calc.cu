// gfb_ewise_calc.cu — BinaryCalc entry points: pick the storage width; each width is its own translation unit
// (gfb_ewise_calc_w{2,4,8}.cu).  One-byte storage cannot hold these fields (m >= 13).
#include "gfb_ewise_api.cuh"

namespace gfb {

gfb_status ewise_binary_calc(BinaryCalc f, int op, int mode, int eb, const void* a, const void* b, uint64_t s_lo, uint64_t,
                             void* out, size_t n, int* err, cudaStream_t st) {
    switch (eb) {
        case 2: return ewise_binary_calc_w2(f, op, mode, a, b, s_lo, out, n, err, st);
        case 4: return ewise_binary_calc_w4(f, op, mode, a, b, s_lo, out, n, err, st);
        case 8: return ewise_binary_calc_w8(f, op, mode, a, b, s_lo, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

gfb_status ewise_unary_calc(BinaryCalc f, int op, int eb, const void* a, void* out, size_t n, int* err, cudaStream_t st) {
    switch (eb) {
        case 2: return ewise_unary_calc_w2(f, op, a, out, n, err, st);
        case 4: return ewise_unary_calc_w4(f, op, a, out, n, err, st);
        case 8: return ewise_unary_calc_w8(f, op, a, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

gfb_status ewise_power_calc(BinaryCalc f, int eb, const void* a, const int64_t* e, int64_t es, bool scalar_e, void* out,
                            size_t n, int* err, cudaStream_t st) {
    switch (eb) {
        case 2: return ewise_power_calc_w2(f, a, e, es, scalar_e, out, n, err, st);
        case 4: return ewise_power_calc_w4(f, a, e, es, scalar_e, out, n, err, st);
        case 8: return ewise_power_calc_w8(f, a, e, es, scalar_e, out, n, err, st);
    }
    return GFB_ERR_BAD_ARG;
}

}  // namespace gfb
