// gfb_ewise_m31.cu — element-wise kernels instantiated for PrimeMersenne31 (see gfb_ewise.cu / gfb_ewise_impl.cuh).
#include "gfb_ewise_impl.cuh"

namespace gfb {

GFB_EWISE_DEFINE(m31, PrimeMersenne31)

}  // namespace gfb
