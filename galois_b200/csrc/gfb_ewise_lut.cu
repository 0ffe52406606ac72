// gfb_ewise_lut.cu — element-wise kernels instantiated for BinaryLut (see gfb_ewise.cu / gfb_ewise_impl.cuh).
#include "gfb_ewise_impl.cuh"

namespace gfb {

GFB_EWISE_DEFINE(lut, BinaryLut)

}  // namespace gfb
