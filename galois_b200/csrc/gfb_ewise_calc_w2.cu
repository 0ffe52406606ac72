// gfb_ewise_calc_w2.cu — element-wise kernels for BinaryCalc (GF(2^m), 13 <= m <= 63) on 2-byte storage.
#include "gfb_ewise_impl.cuh"

namespace gfb {

GFB_EWISE_DEFINE_W(calc_w2, BinaryCalc, uint16_t)

}  // namespace gfb
