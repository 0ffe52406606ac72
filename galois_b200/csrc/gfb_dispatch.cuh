// gfb_dispatch.cuh — (field implementation, storage width) -> template instantiation.
//
// `dispatch_field(f, elem_bytes, fn)` calls `fn(functor, Tag<T>{})` with the device functor of the field
// (gfb_common.cuh) and the storage type of the caller's arrays; fn is a generic lambda that launches the
// kernel templated on both.  Returns GFB_ERR_BAD_ARG when the width cannot hold the field's elements.
#pragma once
#include "gfb_common.cuh"

namespace gfb {

template <typename T>
struct Tag {
    using type = T;
};

inline bool storage_width_ok(const gfb_field* f, int eb) {
    if (f->impl == IMPL_BIN_128) return eb == 16;
    if (eb != 1 && eb != 2 && eb != 4 && eb != 8) return false;
    if (eb == 8) return true;
    if (f->kind == GFB_PRIME) return (f->order_lo - 1) < (1ull << (8 * eb));
    return f->degree <= 8 * eb;
}

template <typename F, typename Fn>
inline gfb_status dispatch_width(const F& functor, int eb, Fn&& fn) {
    switch (eb) {
        case 1: return fn(functor, Tag<uint8_t>{});
        case 2: return fn(functor, Tag<uint16_t>{});
        case 4: return fn(functor, Tag<uint32_t>{});
        case 8: return fn(functor, Tag<unsigned long long>{});
    }
    return GFB_ERR_BAD_ARG;
}

// Fields whose elements fit 64 bits (everything but IMPL_BIN_128, which callers handle or reject themselves).
template <typename Fn>
inline gfb_status dispatch_field(const gfb_field* f, int eb, Fn&& fn) {
    if (!f || !storage_width_ok(f, eb)) return GFB_ERR_BAD_ARG;
    switch (f->impl) {
        case IMPL_PRIME_M31: return dispatch_width(f->prime_m31(), eb, fn);
        case IMPL_PRIME_BARRETT: return dispatch_width(f->prime_barrett(), eb, fn);
        case IMPL_BIN_LUT: return dispatch_width(f->bin_lut(), eb, fn);
        case IMPL_BIN_CALC: return dispatch_width(f->bin_calc(), eb, fn);
        default: break;
    }
    return GFB_ERR_UNSUPPORTED;
}

template <typename F>
inline size_t field_smem_bytes(const F& f) {
    if constexpr (F::kNeedsSmem)
        return (f.smem_bytes() + 15) & ~(size_t)15;
    else
        return 0;
}

}  // namespace gfb
