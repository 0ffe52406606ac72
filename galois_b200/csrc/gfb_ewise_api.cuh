// gfb_ewise_api.cuh — per-functor element-wise launchers (defined in gfb_ewise_<impl>.cu, one translation unit per
// field functor so that the template instantiations compile in parallel) and the operand modes they share with the
// dispatcher in gfb_ewise.cu.
#pragma once
#include "gfb_common.cuh"

namespace gfb {

enum Mode { MODE_AA = 0, MODE_AS = 1, MODE_SA = 2 };  // array∘array, array∘scalar, scalar∘array

#define GFB_EWISE_DECLARE(NAME, FUNCTOR)                                                                            \
    gfb_status ewise_binary_##NAME(FUNCTOR f, int op, int mode, int eb, const void* a, const void* b, uint64_t s_lo, \
                                   uint64_t s_hi, void* out, size_t n, int* err, cudaStream_t st);                  \
    gfb_status ewise_unary_##NAME(FUNCTOR f, int op, int eb, const void* a, void* out, size_t n, int* err,          \
                                  cudaStream_t st);                                                                  \
    gfb_status ewise_power_##NAME(FUNCTOR f, int eb, const void* a, const int64_t* e, int64_t es, bool scalar_e,     \
                                  void* out, size_t n, int* err, cudaStream_t st);

// one storage width of one functor (BinaryCalc is split this way: its carry-less multiply makes each width a long compile)
#define GFB_EWISE_DECLARE_W(NAME, FUNCTOR)                                                                          \
    gfb_status ewise_binary_##NAME(FUNCTOR f, int op, int mode, const void* a, const void* b, uint64_t s, void* out, \
                                   size_t n, int* err, cudaStream_t st);                                             \
    gfb_status ewise_unary_##NAME(FUNCTOR f, int op, const void* a, void* out, size_t n, int* err, cudaStream_t st);  \
    gfb_status ewise_power_##NAME(FUNCTOR f, const void* a, const int64_t* e, int64_t es, bool scalar_e, void* out,   \
                                  size_t n, int* err, cudaStream_t st);

GFB_EWISE_DECLARE_W(calc_w2, BinaryCalc)
GFB_EWISE_DECLARE_W(calc_w4, BinaryCalc)
GFB_EWISE_DECLARE_W(calc_w8, BinaryCalc)

GFB_EWISE_DECLARE(m31, PrimeMersenne31)
GFB_EWISE_DECLARE(barrett, PrimeBarrett)
GFB_EWISE_DECLARE(lut, BinaryLut)
GFB_EWISE_DECLARE(calc, BinaryCalc)
GFB_EWISE_DECLARE(b128, Binary128)

// byte-sized GF(2^m) multiply / divide through the 64 KiB product table (gfb_ewise_tab8.cu)
gfb_status ensure_tab8(gfb_field* f, cudaStream_t st);  // builds f->d_tab8 (products, then quotients) on first use
bool tab8_applies(const gfb_field* f, int op, int eb, const void* a, const void* b, const void* out);
gfb_status ewise_binary_tab8(gfb_field* f, int op, int mode, const void* a, const void* b, uint64_t s, void* out, size_t n,
                             cudaStream_t st);

}  // namespace gfb
