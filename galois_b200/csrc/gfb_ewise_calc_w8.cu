// gfb_ewise_calc_w8.cu — element-wise kernels for BinaryCalc (GF(2^m), 13 <= m <= 63) on 8-byte storage.
#include "gfb_ewise_impl.cuh"

namespace gfb {

GFB_EWISE_DEFINE_W(calc_w8, BinaryCalc, unsigned long long)

}  // namespace gfb
