// gfb_ewise_barrett.cu — element-wise kernels instantiated for PrimeBarrett (see gfb_ewise.cu / gfb_ewise_impl.cuh).
#include "gfb_ewise_impl.cuh"

namespace gfb {

GFB_EWISE_DEFINE(barrett, PrimeBarrett)

}  // namespace gfb
