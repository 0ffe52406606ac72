// gfb_ewise_b128.cu — element-wise kernels instantiated for Binary128 (GF(2^m), 64 <= m <= 128, 16-byte elements).
#include "gfb_ewise_impl.cuh"

namespace gfb {

gfb_status ewise_binary_b128(Binary128 f, int op, int mode, int, const void* a, const void* b, uint64_t s_lo, uint64_t s_hi,
                             void* out, size_t n, int* err, cudaStream_t st) {
    return launch_binary_mode<u128, Binary128>(f, op, mode, a, b, u128{s_lo, s_hi}, out, n, err, st);
}

gfb_status ewise_unary_b128(Binary128 f, int op, int, const void* a, void* out, size_t n, int* err, cudaStream_t st) {
    return launch_unary_op<u128, Binary128>(f, op, a, out, n, err, st);
}

gfb_status ewise_power_b128(Binary128 f, int, const void* a, const int64_t* e, int64_t es, bool scalar_e, void* out, size_t n,
                            int* err, cudaStream_t st) {
    return launch_power_t<u128, Binary128, true>(f, a, e, es, scalar_e, out, n, err, st);
}

}  // namespace gfb
