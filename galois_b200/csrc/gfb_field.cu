// gfb_field.cu — library-level entry points, field handles, lookup-table construction.
//
// Table construction restates FieldArray `_build_lookup_tables`
// (/root/reference/src/galois/_domains/_lookup.py:319-371): EXP[i] = alpha^i for i < order,
// LOG[EXP[i]] = i for 1 <= i < order-1, LOG[1] = 0, and EXP doubled so that
// EXP[LOG[a] + LOG[b]] needs no modulo (:370-371).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gfb_common.cuh"

namespace gfb {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* what, cudaError_t err) {
    snprintf(g_last_error, sizeof(g_last_error), "%s failed: %s (%s)", what, cudaGetErrorName(err),
             cudaGetErrorString(err));
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// host carry-less multiply with on-the-fly reduction, m <= 16 (table build only)
static uint32_t host_mul_binary(uint32_t a, uint32_t b, int m, uint32_t poly) {
    uint32_t c = 0;
    const uint32_t order = 1u << m;
    while (b) {
        if (b & 1u) c ^= a;
        b >>= 1;
        a <<= 1;
        if (a >= order) a ^= poly;
    }
    return c;
}

}  // namespace gfb

namespace gfb {
// deterministic Miller-Rabin for n < 2^32 (bases 2, 7, 61)
static uint32_t mulmod_u32(uint32_t a, uint32_t b, uint32_t n) { return (uint32_t)((uint64_t)a * b % n); }
static uint32_t powmod_u32(uint32_t a, uint32_t e, uint32_t n) {
    uint32_t r = 1;
    while (e) {
        if (e & 1) r = mulmod_u32(r, a, n);
        a = mulmod_u32(a, a, n);
        e >>= 1;
    }
    return r;
}
static bool is_prime_u32(uint32_t n) {
    if (n < 2) return false;
    for (uint32_t q : {2u, 3u, 5u, 7u, 11u, 13u, 17u, 19u, 23u, 29u, 31u, 37u, 61u}) {
        if (n == q) return true;
        if (n % q == 0) return false;
    }
    uint32_t d = n - 1;
    int r = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++r;
    }
    for (uint32_t a : {2u, 7u, 61u}) {
        uint32_t x = powmod_u32(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < r; ++i) {
            x = mulmod_u32(x, x, n);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}
}  // namespace gfb

extern "C" {

int gfb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* gfb_last_error(void) { return gfb::g_last_error; }

const char* gfb_version(void) { return "galois_b200 0.1 (sm_100a)"; }

void* gfb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 1;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        gfb::set_last_error("cudaHostAlloc", e);
        return nullptr;
    }
    return p;
}

void gfb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

gfb_status gfb_field_create(gfb_field** out, int kind, uint64_t order_lo, int degree, uint64_t poly_lo,
                            uint64_t poly_hi, uint64_t primitive_element) {
    if (!out) return GFB_ERR_BAD_ARG;
    *out = nullptr;
    gfb_field* f = (gfb_field*)calloc(1, sizeof(gfb_field));
    if (!f) return GFB_ERR_BAD_ARG;
    f->kind = kind;
    f->primitive_element = primitive_element;
    if (kind == GFB_PRIME) {
        if (order_lo < 2 || order_lo >= (1ull << 32)) {
            free(f);
            return order_lo >= (1ull << 32) ? GFB_ERR_UNSUPPORTED : GFB_ERR_BAD_ARG;
        }
        if (!gfb::is_prime_u32((uint32_t)order_lo)) {
            // even p > 2 has no p^-1 mod 2^32 and a composite p breaks Fermat inversion: refuse instead of
            // returning garbage from mul / inv (the Python factory validates too, but this is the boundary)
            free(f);
            return GFB_ERR_BAD_ARG;
        }
        f->degree = 1;
        f->order_lo = order_lo;
        f->p = (uint32_t)order_lo;
        // floor(2^64 / p): p never divides 2^64 for odd p; for p = 2, (2^64-1)/2 + ... handle generally
        unsigned __int128 one = (unsigned __int128)1 << 64;
        f->mu = (unsigned long long)(one / order_lo);
        if (order_lo == 2) f->mu = 0x8000000000000000ull;  // exact 2^63
        // Montgomery constants: p^-1 mod 2^32 by Newton iteration (odd p), 2^64 mod p
        uint32_t inv = f->p;  // correct to 3 bits
        for (int i = 0; i < 5; ++i) inv *= 2u - f->p * inv;
        f->pinv = inv;
        f->r2 = (uint32_t)(one % order_lo);
        f->impl = (order_lo == 0x7fffffffull) ? gfb::IMPL_PRIME_M31 : gfb::IMPL_PRIME_BARRETT;
    } else if (kind == GFB_BINARY) {
        if (degree < 1 || degree > 128) {
            free(f);
            return GFB_ERR_BAD_ARG;
        }
        f->degree = degree;
        f->order_lo = degree < 64 ? (1ull << degree) : 0ull;
        f->poly_lo = poly_lo;
        f->poly_hi = poly_hi;
        if (degree < 64 && ((poly_lo >> degree) != 1ull || poly_hi != 0)) {
            // the x^m term must be present and nothing above it
            free(f);
            return GFB_ERR_BAD_ARG;
        }
        if (degree <= 12)
            f->impl = gfb::IMPL_BIN_LUT;
        else if (degree <= 63)
            f->impl = gfb::IMPL_BIN_CALC;
        else
            f->impl = gfb::IMPL_BIN_128;
    } else {
        free(f);
        return GFB_ERR_BAD_ARG;
    }

    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        gfb::set_last_error("cudaGetDevice", e);
        free(f);
        return GFB_ERR_CUDA;
    }
    f->device = dev;

    e = cudaHostAlloc((void**)&f->h_err, sizeof(int), cudaHostAllocMapped);
    if (e != cudaSuccess) {
        gfb::set_last_error("cudaHostAlloc(mapped)", e);
        free(f);
        return GFB_ERR_CUDA;
    }
    *f->h_err = 0;
    e = cudaHostGetDevicePointer((void**)&f->d_err, f->h_err, 0);
    if (e != cudaSuccess) {
        gfb::set_last_error("cudaHostGetDevicePointer", e);
        cudaFreeHost(f->h_err);
        free(f);
        return GFB_ERR_CUDA;
    }

    if (kind == GFB_BINARY && degree <= 16) {
        const uint32_t order = 1u << degree;
        const uint32_t poly = (uint32_t)poly_lo;
        const uint32_t alpha = (uint32_t)primitive_element;
        if (alpha == 0 || alpha >= order) {
            gfb_field_destroy(f);
            return GFB_ERR_BAD_ARG;
        }
        f->h_exp = (uint16_t*)calloc(2u * order, sizeof(uint16_t));
        f->h_log = (uint16_t*)calloc(order, sizeof(uint16_t));
        uint32_t element = 1;
        f->h_exp[0] = 1;
        bool ok = true;
        if (degree == 1) {
            f->h_exp[1] = 1;
        }
        for (uint32_t i = 1; i < order; ++i) {
            element = gfb::host_mul_binary(element, alpha, degree, poly);
            f->h_exp[i] = (uint16_t)element;
            if (i < order - 1) {
                if (element == 1u || (element != 1u && f->h_log[element] != 0)) ok = false;  // not a generator
                f->h_log[element] = (uint16_t)i;
            }
        }
        if (f->h_exp[order - 1] != 1) ok = false;
        if (!ok) {
            gfb_field_destroy(f);
            return GFB_ERR_BAD_ARG;
        }
        for (uint32_t i = 0; i < order; ++i) f->h_exp[order + i] = f->h_exp[1 + i];
        f->h_exp[2u * order - 1] = f->h_exp[order];  // EXP[1 + order - 1] = EXP[order] (wrap)
        e = cudaMalloc((void**)&f->d_exp, 2u * order * sizeof(uint16_t));
        if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_log, order * sizeof(uint16_t));
        if (e == cudaSuccess)
            e = cudaMemcpy(f->d_exp, f->h_exp, 2u * order * sizeof(uint16_t), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(f->d_log, f->h_log, order * sizeof(uint16_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            gfb::set_last_error("field table upload", e);
            gfb_field_destroy(f);
            return GFB_ERR_CUDA;
        }
    }
    *out = f;
    return GFB_OK;
}

void gfb_field_destroy(gfb_field* f) {
    if (!f) return;
    gfb::conv_cache_drop(f);
    if (f->d_tab8) cudaFree(f->d_tab8);
    if (f->d_exp) cudaFree(f->d_exp);
    if (f->d_log) cudaFree(f->d_log);
    if (f->h_err) cudaFreeHost(f->h_err);
    free(f->h_exp);
    free(f->h_log);
    free(f);
}

int gfb_field_uses_tables(const gfb_field* f) { return f && f->impl == gfb::IMPL_BIN_LUT; }

gfb_status gfb_take_error(gfb_field* f, void* stream) {
    if (!f) return GFB_ERR_BAD_ARG;
    GFB_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    int v = *(volatile int*)f->h_err;
    *(volatile int*)f->h_err = 0;
    return v ? GFB_ERR_ZERO_DIVISION : GFB_OK;
}

}  // extern "C"
