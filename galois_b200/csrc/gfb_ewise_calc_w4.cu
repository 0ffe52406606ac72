// gfb_ewise_calc_w4.cu — element-wise kernels for BinaryCalc (GF(2^m), 13 <= m <= 63) on 4-byte storage.
#include "gfb_ewise_impl.cuh"

namespace gfb {

GFB_EWISE_DEFINE_W(calc_w4, BinaryCalc, uint32_t)

}  // namespace gfb
